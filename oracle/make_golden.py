"""TEST INFRASTRUCTURE ONLY - generates ``tests/golden/*.npz`` by running the REAL reference.

Run here (the container that has /root/reference):  ``python -m oracle.make_golden``

The reference has no golden vectors of its own, so the fixtures are the outputs of
its own functions, imported read-only through ``oracle/ref_loader.py``, on seeded synthetic inputs.
Inputs are NOT stored when they can be regenerated from a seed (``oracle.dh_oracle.synthetic_scene``);
large outputs are stored as SHA-256 digests plus the small arrays needed to localise a mismatch.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dh_oracle as O          # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (scene kwargs, angle, axis, translation, use_input_depth_normalization)
PC_CASES = {
    "cfg1":      (dict(S=512, seed=0), 30.0, (0.0, 1.0, 0.0), (0.3, 0.0, 0.2), False),
    "neg60":     (dict(S=512, seed=1, cx=300.0, cy=220.0, radius=90.0), -60.0, (0.0, 1.0, 0.0), (-1.0, 0.0, 0.5), False),
    "occl90":    (dict(S=512, seed=2, radius=150.0), 90.0, (0.0, 1.0, 0.0), (1.5, 0.0, 1.0), False),
    "zties45":   (dict(S=512, seed=3, quantize=0.1), 45.0, (0.0, 1.0, 0.0), (0.2, 0.1, 0.0), False),
    "xaxis20":   (dict(S=512, seed=4, radius=100.0), 20.0, (1.0, 0.0, 0.0), (0.0, -0.2, 0.3), False),
    "identity":  (dict(S=512, seed=5), 0.0, (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), False),
    "cfg1_norm": (dict(S=512, seed=0), 30.0, (0.0, 1.0, 0.0), (0.3, 0.0, 0.2), True),
    "zaxis_all_offscreen": (dict(S=512, seed=6, radius=60.0), 60.0, (0.0, 0.0, 1.0), (-4.0, 0.0, -1.5), False),
    "axis_scaled": (dict(S=512, seed=7, radius=80.0), -35.0, (0.0, 2.5, 0.0), (0.1, 0.05, -0.3), False),
    # smooth surfaces, irregular mask with a hole, a thin bar and a detached blob (closer to the reference's fixtures)
    "smooth25":  (dict(S=512, seed=21, kind="smooth"), 25.0, (0.0, 1.0, 0.0), (0.25, 0.0, 0.1), False),
    "smooth_m50": (dict(S=512, seed=22, kind="smooth"), -50.0, (0.0, 1.0, 0.0), (-0.4, 0.05, 0.3), True),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_pc_intermediates(ref, depth, bg, mask, K, angle, axis, t):
    """Re-run the reference's own building blocks the way transform_depth_pc (depth_transform.py:226-281)
    assembles them, to obtain the intermediates the public function does not return."""
    dt = ref.depth_transform
    S = mask.shape[-1]
    td, tb = torch.from_numpy(depth)[None, None], torch.from_numpy(bg)[None, None]
    bg_pts = dt.depth_to_world_coords(tb, intrinsics=K)
    pts = dt.depth_to_world_coords(td, intrinsics=K)
    tt = torch.tensor(t, dtype=torch.float32)
    pts2, mod = dt.transform_point_cloud(points=pts.numpy(), axis=np.asarray(axis, np.float32), angle_degrees=angle,
                                         x=tt[0].item(), y=tt[1].item(), z=tt[2].item(), mask=mask)
    rb = bg_pts.numpy().reshape(S * S, 3)
    rp = pts2.reshape(S * S, 3)
    ids = np.where(mod)[0]
    allp = np.vstack([rb, rp[ids]])
    pm = np.zeros(len(allp), np.uint8)
    pm[S * S:] = 1
    depth_map, depth_mask, tx, ty, vis = dt.points_to_depth(
        points=torch.from_numpy(allp), intrinsics=K, output_size=(S, S), point_mask=torch.from_numpy(pm))
    return allp, depth_map[0, 0].numpy(), depth_mask, tx, ty, vis


def make_pc_goldens(ref):
    K = ref.get_depth_intrinsics()
    out, meta = {}, {}
    for name, (scene, angle, axis, t, norm) in PC_CASES.items():
        depth, bg, mask = O.synthetic_scene(**scene)
        S = scene["S"]
        if S != 512:
            raise ValueError("the reference's pc path is hard-wired to 512 (depth_transform.py:531)")
        td, tb, tm = (torch.from_numpy(a)[None, None] for a in (depth, bg, mask))
        disp, corr = ref.depth_transform.transform_depth_pc(
            td, tb, tm, K, rot_angle=angle, rot_axis=torch.tensor(axis, dtype=torch.float32),
            translation=torch.tensor(t, dtype=torch.float32), use_input_depth_normalization=norm)
        allp, depth_map, depth_mask, tx, ty, vis = ref_pc_intermediates(ref, depth, bg, mask, K, angle, axis, t)
        disp = disp[0, 0].numpy()
        meta[name] = dict(scene=scene, angle=angle, axis=list(axis), translation=list(t), norm=norm,
                          n_corr=int(corr.shape[0]), n_fg=int(mask.sum()),
                          sha_points=sha(allp), sha_depth_map=sha(depth_map), sha_disparity=sha(disp),
                          sha_visible=sha(np.packbits(vis)), sha_inputs=sha(depth) + sha(bg) + sha(mask))
        out[f"{name}/corr"] = corr.numpy().astype(np.int16)
        out[f"{name}/target_mask"] = np.packbits(depth_mask)
        out[f"{name}/visible_count"] = np.int64(vis.sum())
        out[f"{name}/depth_map_rows"] = depth_map[::32].copy()
        out[f"{name}/disparity_rows"] = disp[::32].copy()
        out[f"{name}/centroid_points_head"] = allp[S * S:S * S + 64].copy()
        print(f"[golden] {name}: n_fg={meta[name]['n_fg']} n_corr={meta[name]['n_corr']}")
    # empty mask branch (depth_transform.py:203-216)
    depth, bg, mask = O.synthetic_scene(S=512, seed=8)
    disp, corr = ref.depth_transform.transform_depth_pc(
        torch.from_numpy(depth)[None, None], torch.from_numpy(bg)[None, None],
        torch.zeros(1, 1, 512, 512), K, rot_angle=10.0)
    meta["empty_mask"] = dict(scene=dict(S=512, seed=8), n_corr=int(corr.shape[0]), sha_disparity=sha(disp[0, 0].numpy()),
                              corr_shape=list(corr.shape), corr_dtype=str(corr.dtype))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pc_transform.npz"), **out)
    with open(os.path.join(GOLDEN_DIR, "pc_transform.json"), "w") as f:
        json.dump(meta, f, indent=1)


def make_small_goldens(ref):
    """Resolution-generic pieces on small inputs; inputs are stored (they are tiny)."""
    dt = ref.depth_transform
    K = ref.get_depth_intrinsics()
    rng = np.random.default_rng(1234)
    out = {}
    # depth_to_world_coords on square and non-square maps
    for tag, (H, W) in {"sq": (64, 64), "wide": (48, 80), "tall": (70, 33)}.items():
        d = (1.5 + 3 * rng.random((H, W))).astype(np.float32)
        out[f"unproj_{tag}/depth"] = d
        out[f"unproj_{tag}/points"] = dt.depth_to_world_coords(torch.from_numpy(d)[None, None], K).numpy()
    # points_to_depth with adversarial points: exact z ties, negative z, far off-screen, shared pixels
    for tag, (H, W, N) in {"a": (40, 40, 6000), "b": (24, 56, 5000)}.items():
        pts = np.empty((N, 3))
        pts[:, 2] = np.round(rng.uniform(0.5, 3.0, N), 1)           # many exact ties
        pts[:, 0] = rng.uniform(-3.0, 3.0, N)
        pts[:, 1] = rng.uniform(-3.0, 3.0, N)
        neg = rng.random(N) < 0.03
        pts[neg, 2] = -pts[neg, 2]
        pm = (rng.random(N) < 0.4).astype(np.uint8)
        dm, mk, tx, ty, vis = dt.points_to_depth(torch.from_numpy(pts), K, (H, W), point_mask=torch.from_numpy(pm))
        out[f"p2d_{tag}/points"] = pts
        out[f"p2d_{tag}/point_mask"] = pm
        out[f"p2d_{tag}/size"] = np.array([H, W])
        out[f"p2d_{tag}/depth_map"] = dm[0, 0].numpy()
        out[f"p2d_{tag}/depth_mask"] = mk
        out[f"p2d_{tag}/tx"] = tx
        out[f"p2d_{tag}/ty"] = ty
        out[f"p2d_{tag}/visible"] = vis
    # transform_points (mesh-mode / webapp variant, torch fp32), depth_transform.py:439-459
    p = rng.normal(size=(500, 3)).astype(np.float32)
    out["tp/points"] = p
    out["tp/out"] = dt.transform_points(torch.from_numpy(p), rot_angle=torch.tensor(25.0), rot_axis=torch.tensor([0.0, 1.0, 0.0]),
                                        translation=torch.tensor([0.1, -0.2, 0.3])).numpy()
    # normalize_depth
    x = rng.random((1, 1, 16, 16)).astype(np.float32)
    out["nd/x"] = x
    out["nd/y"] = dt.normalize_depth(torch.from_numpy(x)).numpy()
    # process_correspondences on the cfg1 correspondences and a 1024 variant
    g = np.load(os.path.join(GOLDEN_DIR, "pc_transform.npz"))
    corr = g["cfg1/corr"].astype(np.int64)
    for tag, (c, res, er) in {"e0": (corr, 512, 0), "e5": (corr, 512, 5), "e15": (corr, 512, 15),
                              "r1024": (corr * 2 + rng.integers(0, 2, corr.shape), 1024, 0),
                              "oob": (np.concatenate([corr[:200], [[5, 5, 600, 3], [7, 7, -1, 9]]]), 512, 1)}.items():
        pc = ref.process_correspondences(torch.from_numpy(np.asarray(c, dtype=np.int64)), res, er)
        out[f"pcorr_{tag}/corr"] = np.asarray(c, dtype=np.int64).astype(np.int16)
        out[f"pcorr_{tag}/res_er"] = np.array([res, er])
        for k, v in pc.items():
            out[f"pcorr_{tag}/{k}"] = np.asarray(v).astype(np.int16)
    # losses fwd + autograd grads, small channel counts, both native sizes, both bg types
    pc = ref.process_correspondences(torch.from_numpy(corr), 512, 0)
    for tag, (C, h) in {"c6h64": (6, 64), "c5h32": (5, 32), "c3h16": (3, 16)}.items():
        cur = torch.from_numpy(rng.normal(size=(C, h, h)).astype(np.float32)).requires_grad_(True)
        orig = torch.from_numpy(rng.normal(size=(C, h, h)).astype(np.float32))
        out[f"loss_{tag}/cur"] = cur.detach().numpy()
        out[f"loss_{tag}/orig"] = orig.numpy()
        lf = ref.losses.compute_foreground_loss(cur, orig, pc, 1, (64, 64))
        out[f"loss_{tag}/fg"] = lf.detach().numpy()
        out[f"loss_{tag}/fg_grad"] = torch.autograd.grad(lf, cur)[0].numpy()
        for lt in ("global_avg", "local_avg"):
            lb = ref.losses.compute_background_loss(cur, orig, pc, 1, (64, 64), loss_type=lt)
            out[f"loss_{tag}/bg_{lt}"] = lb.detach().numpy()
            out[f"loss_{tag}/bg_{lt}_grad"] = torch.autograd.grad(lb, cur)[0].numpy()
    # transform_point_cloud (hard-wired to 512 x 512 in the reference, depth_transform.py:531)
    depth, bg, mask = O.synthetic_scene(S=512, seed=9, radius=70.0)
    pts = dt.depth_to_world_coords(torch.from_numpy(depth)[None, None], K).numpy()
    rot, mod = dt.transform_point_cloud(pts, np.array([0.0, 1.0, 0.0], np.float32), 33.0, 0.25, -0.5, 0.125, mask)
    out["tpc/sha"] = np.frombuffer(bytes.fromhex(sha(rot)), dtype=np.uint8)
    out["tpc/rows"] = rot[::64].copy()
    out["tpc/mod_count"] = np.int64(mod.sum())
    # solve_laplacian_depth + the set_foreground recipe (diffusion_handles.py:105-108), resolution generic
    import scipy.ndimage
    for tag, (S, it) in {"s96": (96, 4), "s160": (160, 15)}.items():
        depth, bg, mask = O.synthetic_scene(S=S, seed=12, radius=S / 6)
        dil = scipy.ndimage.binary_dilation(mask.astype(bool), iterations=it)
        sol = ref.utils.solve_laplacian_depth(depth, bg, dil)
        out[f"sld_{tag}/S_it"] = np.array([S, it])
        out[f"sld_{tag}/dilated"] = np.packbits(dil)
        out[f"sld_{tag}/solution"] = sol
    # local-average losses with patch_size > 1 (losses.py:62-77): odd and even patches, native and resized maps,
    # separate generator so that the fixtures above keep their values
    rng2 = np.random.default_rng(77)
    for tag, (C, h, patch) in {"p3c4h64": (4, 64, 3), "p5c3h32": (3, 32, 5), "p2c3h16": (3, 16, 2), "p4c2h64": (2, 64, 4)}.items():
        cur = torch.from_numpy(rng2.normal(size=(C, h, h)).astype(np.float32)).requires_grad_(True)
        orig = torch.from_numpy(rng2.normal(size=(C, h, h)).astype(np.float32))
        out[f"ploss_{tag}/cur"] = cur.detach().numpy()
        out[f"ploss_{tag}/orig"] = orig.numpy()
        out[f"ploss_{tag}/patch"] = np.int64(patch)
        lf = ref.losses.compute_foreground_loss(cur, orig, pc, patch, (64, 64))
        out[f"ploss_{tag}/fg"] = lf.detach().numpy()
        out[f"ploss_{tag}/fg_grad"] = torch.autograd.grad(lf, cur)[0].numpy()
        lb = ref.losses.compute_background_loss(cur, orig, pc, patch, (64, 64), loss_type="local_avg")
        out[f"ploss_{tag}/bg_local_avg"] = lb.detach().numpy()
        out[f"ploss_{tag}/bg_local_avg_grad"] = torch.autograd.grad(lb, cur)[0].numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "small_cases.npz"), **out)
    print(f"[golden] small cases: {len(out)} arrays")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = load_reference()
    make_pc_goldens(ref)
    make_small_goldens(ref)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
