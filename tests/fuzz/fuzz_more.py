"""More randomised sweeps against the oracle:
  batch    EditEngine batches of 2..6 random edits (mixed mask sizes, some empty) - per-edit winners, correspondences,
           masks and raw disparity bit-exact (exercises the batched kernels' per-edit offsets)
  poisson  edits with the Poisson hole fill, |disparity - oracle (SuperLU)| <= 1e-3 on the 0..255 scale
  raster   the triangle rasteriser on random depth-map meshes (sizes 16..48, random rigid moves): pix_to_face, zbuf,
           barycentrics bit-exact against the NumPy restatement of pytorch3d's semantics
python tests/fuzz/fuzz_more.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dh_oracle as O                                        # noqa: E402
from diffusionhandles_b200 import depth_transform as dt                 # noqa: E402
from diffusionhandles_b200.engine import EditEngine, make_rigid          # noqa: E402
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser   # noqa: E402
from diffusionhandles_b200.renderer import Camera, MeshRenderer, SplatRendererArgs    # noqa: E402

dev = torch.device("cuda:0")
K = GuidedStableDiffuser.get_depth_intrinsics()
K_NP = K.numpy()


def random_edit(rng, S):
    kind = "smooth" if rng.random() < 0.4 else "disc"
    seed = int(rng.integers(0, 10_000))
    scene = dict(S=S, seed=seed, kind="smooth") if kind == "smooth" else dict(
        S=S, seed=seed, cx=float(rng.uniform(0.2, 0.8) * S), cy=float(rng.uniform(0.2, 0.8) * S), radius=float(rng.uniform(0.03, 0.4) * S))
    depth, bg, mask = O.synthetic_scene(**scene)
    if rng.random() < 0.15:
        mask = np.zeros_like(mask)
    angle = float(rng.uniform(-90, 90))
    axis = [(0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0)][rng.integers(0, 3)]
    t = tuple(float(v) for v in rng.normal(size=3) * 0.3)
    return depth, bg, mask, angle, axis, t


def fuzz_batch(rng) -> bool:
    S = int(rng.choice([64, 128, 256]))
    B = int(rng.integers(2, 7))
    edits = [random_edit(rng, S) for _ in range(B)]
    eng = EditEngine(dev, B, S, S)
    td = torch.from_numpy(np.stack([e[0] for e in edits])).to(dev)
    tb = torch.from_numpy(np.stack([e[1] for e in edits])).to(dev)
    tm = torch.from_numpy(np.stack([e[2] for e in edits])).to(dev)
    norm = bool(rng.random() < 0.3)
    res = eng.run(td, tb, tm, K, [make_rigid(e[3], list(e[4]), list(e[5])) for e in edits], use_input_depth_normalization=norm, poisson=False)
    ok = True
    for i, (depth, bg, mask, angle, axis, t) in enumerate(edits):
        if not mask.any():
            ok &= int(res.n_corr_host[i]) == 0 and int(res.n_fg_host[i]) == 0
            continue
        t32 = tuple(float(np.float32(v)) for v in t)
        o = O.transform_depth_pc(depth, bg, mask, K_NP, angle, axis, t32, use_input_depth_normalization=norm, poisson=False)
        n = int(res.n_corr_host[i])
        ok &= np.array_equal(res.winner[i].cpu().numpy().astype(np.int64), o["winner"])
        ok &= n == o["correspondences"].shape[0] and np.array_equal(res.corr[i, :n].cpu().numpy(), o["correspondences"])
        ok &= np.array_equal(eng.unpack_bits(res.cleaned_bits)[i].cpu().numpy().astype(bool), o["cleaned"])
        ok &= np.array_equal(res.disparity_raw[i].cpu().numpy(), o["disparity_raw"], equal_nan=True)
    return bool(ok)


def fuzz_poisson(rng) -> bool:
    S = int(rng.choice([64, 96, 128, 192]))
    depth, bg, mask, angle, axis, t = random_edit(rng, S)
    if not mask.any():
        return True
    t32 = tuple(float(np.float32(v)) for v in t)
    o = O.transform_depth_pc(depth, bg, mask, K_NP, angle, axis, t32, poisson=True)
    disp, corr = dt.transform_depth_pc(torch.from_numpy(depth).to(dev)[None, None], torch.from_numpy(bg).to(dev)[None, None],
                                       torch.from_numpy(mask).to(dev)[None, None], K, rot_angle=angle,
                                       rot_axis=torch.tensor(axis), translation=torch.tensor(t, dtype=torch.float32))
    return bool(np.array_equal(corr.numpy(), o["correspondences"]) and np.abs(disp[0, 0].cpu().numpy() - o["disparity"]).max() <= 1e-3)


def fuzz_raster(rng) -> bool:
    S = int(rng.choice([16, 24, 32, 48]))
    depth, bg, mask, angle, axis, t = random_edit(rng, S)
    if not mask.any():
        return True
    td, tb = torch.from_numpy(depth).to(dev)[None, None], torch.from_numpy(bg).to(dev)[None, None]
    tm = torch.from_numpy(mask > 0.5).to(dev)[None, None]
    bg_mesh, fg_mesh = dt.depth_to_mesh(tb, K), dt.depth_to_mesh(td, K, mask=tm)
    if fg_mesh.verts.shape[0] == 0:
        return True
    fg_mesh.verts = dt.transform_points(fg_mesh.verts, torch.tensor(angle), torch.tensor(axis), torch.tensor(t, dtype=torch.float32))
    cull = bool(rng.random() < 0.7)
    blur = float(rng.choice([0.0, 1e-5, 1e-3]))
    r = MeshRenderer(['world_position', 'flat_vertex_color'], SplatRendererArgs(device=dev, output_res=(S, S), cull_backfaces=cull, blur_radius=blur))
    r.update_scene({'meshes': [bg_mesh, fg_mesh], 'cameras': [Camera(intrinsics=K)]})
    out = r.render()
    p2f, zbuf, bary = (x.cpu().numpy() for x in r.fragments[0])
    verts = torch.cat([bg_mesh.verts, fg_mesh.verts]).cpu().numpy()
    faces = torch.cat([bg_mesh.faces, fg_mesh.faces + bg_mesh.verts.shape[0]]).cpu().numpy()
    sx, sy = O.fov_scales(float(K_NP[1, 1]), S, S)
    o_p2f, o_z, o_b = O.rasterize_meshes(verts, faces, S, S, sx, sy, blur, cull, True, True)
    ok = np.array_equal(p2f, o_p2f) and np.array_equal(zbuf, o_z) and np.array_equal(bary, o_b)
    ok &= np.array_equal(out['world_position'][0].cpu().numpy(), O.interpolate_face_attributes(verts, faces, o_p2f, o_b))
    return bool(ok)


def run(n_cases: int, seed: int, verbose: bool = True, which=("batch", "poisson", "raster")) -> int:
    rng = np.random.default_rng(seed)
    fns = {"batch": fuzz_batch, "poisson": fuzz_poisson, "raster": fuzz_raster}
    bad = 0
    for case in range(n_cases):
        for name in which:
            if not fns[name](rng):
                bad += 1
                if verbose:
                    print(f"MISMATCH {name} case {case} (seed {seed})", flush=True)
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    t0 = time.time()
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(f"{n} cases x 3 sweeps, {bad} mismatching, {time.time() - t0:.0f} s")
