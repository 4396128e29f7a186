"""TEST INFRASTRUCTURE ONLY - packs the reference's bundled photogen scenes and pins the REAL reference on them.

Run here (the container that has /root/reference):  ``python -m oracle.make_golden_photogen``

Writes
  tests/golden/photogen_inputs.npz   all 20 scenes of /root/reference/test/data/photogen as loaded by the reference's test driver
                                     (test/test_diffusion_handles.py:208-263): depth / bg_depth (the EXR values are fp16-exact, stored
                                     as fp16 - asserted lossless), mask (> 0.5, bit-packed)
  tests/golden/photogen_ref.json     per scene: the transforms (transforms.json), SHA-256 of the reference's set_foreground result
                                     (diffusion_handles.py:105-108: solve_laplacian_depth over the 15 px dilated mask);
                                     per edit (90): n_corr and SHA-256 of the correspondences / disparity returned by the reference's
                                     transform_depth_pc on (depth, set_foreground(bg_depth), mask) - the call transform_foreground makes
  tests/golden/pins.npz              points_to_depth (depth_transform.py:643-747) called DIRECTLY at 1024^2 on the config-5 A/B/C point
                                     sets (SHA-256 of every output), and the StepGuidanceWeightSchedule class
                                     (guided_stable_diffuser.py:622-665) evaluated on the schedule guided_inference builds (:336-373)
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dh_oracle as O                                              # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, load_exr, load_reference          # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PHOTOGEN = os.path.join(REFERENCE_ROOT, "test", "data", "photogen")
CONFIG5 = {"A": (None, 90.0, (1.5, 0.0, 1.0)), "B": (0.1, 90.0, (1.5, 0.0, 1.0)), "C": (None, 60.0, (-2.0, 0.0, -1.5))}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_scene(name: str):
    """The arrays the reference's test driver hands to DiffusionHandles (512^2 fixtures: crop_and_resize is the identity)."""
    import cv2
    d = os.path.join(PHOTOGEN, name)
    depth = load_exr(os.path.join(d, "depth.exr")).astype(np.float32)
    bg = load_exr(os.path.join(d, "bg_depth.exr")).astype(np.float32)
    m = cv2.imread(os.path.join(d, "mask.png"), cv2.IMREAD_UNCHANGED)
    if m.ndim == 3:
        m = m.astype(np.float32).mean(-1)
    mask = (m.astype(np.float32) / 255.0 > 0.5).astype(np.float32)
    with open(os.path.join(d, "transforms.json")) as f:
        tr = json.load(f)
    assert depth.shape == bg.shape == mask.shape == (512, 512), name
    return depth, bg, mask, tr


def make_photogen(ref):
    import scipy.ndimage
    K = ref.get_depth_intrinsics()
    arrays, meta = {}, {}
    scenes = sorted(s for s in os.listdir(PHOTOGEN) if os.path.isdir(os.path.join(PHOTOGEN, s)))
    n_edits = 0
    for s in scenes:
        depth, bg, mask, tr = load_scene(s)
        for tag, a in (("depth", depth), ("bg_depth", bg)):
            h = a.astype(np.float16)
            assert np.array_equal(h.astype(np.float32), a), f"{s}/{tag} is not fp16-exact"
            arrays[f"{s}/{tag}"] = h
        arrays[f"{s}/mask"] = np.packbits(mask.astype(bool))
        t0 = time.time()
        dil = scipy.ndimage.binary_dilation(mask, iterations=15)                 # diffusion_handles.py:105-108
        bg2 = ref.utils.solve_laplacian_depth(depth, bg, dil)
        t_fill = time.time() - t0
        ms = dict(n_fg=int(mask.sum()), n_unknown=int(dil.sum()), sha_set_foreground=sha(bg2), set_foreground_dtype=str(bg2.dtype),
                  ref_seconds_set_foreground=round(t_fill, 3), edits={})
        # a few rows inside the hole, so that a tolerance comparison (the CUDA fill is CG, the reference SuperLU) needs no SuperLU
        ys = np.nonzero(dil.any(1))[0]
        rows = ys[:: max(1, len(ys) // 6)][:6]
        arrays[f"{s}/set_foreground_rows_idx"] = rows.astype(np.int16)
        arrays[f"{s}/set_foreground_rows"] = bg2[rows].astype(np.float32)
        for name, t in tr.items():
            t0 = time.time()
            disp, corr = ref.depth_transform.transform_depth_pc(
                torch.from_numpy(depth)[None, None], torch.from_numpy(bg2.astype(np.float32))[None, None],
                torch.from_numpy(mask)[None, None], K, rot_angle=t["rotation_angle"],
                rot_axis=torch.tensor(t["rotation_axis"], dtype=torch.float32),
                translation=torch.tensor(t["translation"], dtype=torch.float32))
            ms["edits"][name] = dict(rotation_angle=t["rotation_angle"], rotation_axis=t["rotation_axis"], translation=t["translation"],
                                     n_corr=int(corr.shape[0]), sha_corr=sha(corr.numpy().astype(np.int64)),
                                     sha_disparity=sha(disp[0, 0].numpy()), ref_seconds=round(time.time() - t0, 3))
            n_edits += 1
        meta[s] = ms
        print(f"[photogen] {s}: n_fg={ms['n_fg']} unknowns={ms['n_unknown']} fill {t_fill:.2f}s, {len(tr)} edits", flush=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "photogen_inputs.npz"), **arrays)
    with open(os.path.join(GOLDEN_DIR, "photogen_ref.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"[photogen] {len(scenes)} scenes, {n_edits} edits")


def config5_points(case: str):
    """The point set of SURVEY.md 8(d) config 5 (1024^2; the reference's transform_point_cloud is hard-wired to 512, so the
    points come from the oracle's resolution-generic restatement - only points_to_depth is pinned here)."""
    q, angle, t = CONFIG5[case]
    S = 1024
    depth, bg, mask = O.synthetic_scene(S, 0, cx=512.0, cy=560.0, radius=300.0, quantize=q)
    o = O.transform_depth_pc(depth, bg, mask, O.get_depth_intrinsics(), angle, (0, 1, 0), tuple(float(np.float32(v)) for v in t),
                             poisson=False)
    return o, S


def make_pins(ref):
    out = {}
    K = ref.get_depth_intrinsics()
    for case in CONFIG5:
        o, S = config5_points(case)
        pm = (np.arange(len(o["points"])) >= S * S).astype(np.uint8)
        t0 = time.time()
        dm, mk, tx, ty, vis = ref.depth_transform.points_to_depth(torch.from_numpy(o["points"]), K, (S, S), point_mask=torch.from_numpy(pm))
        print(f"[pins] config5{case}: reference points_to_depth at 1024^2 took {time.time() - t0:.1f}s, visible={int(vis.sum())}", flush=True)
        for k, v in (("depth_map", dm[0, 0].numpy()), ("depth_mask", np.packbits(mk)), ("tx", np.asarray(tx, np.int64)),
                     ("ty", np.asarray(ty, np.int64)), ("visible", np.packbits(vis))):
            out[f"p2d1024_{case}/sha_{k}"] = np.frombuffer(bytes.fromhex(sha(v)), dtype=np.uint8)
        out[f"p2d1024_{case}/sha_points"] = np.frombuffer(bytes.fromhex(sha(o["points"])), dtype=np.uint8)
        out[f"p2d1024_{case}/n_visible"] = np.int64(vis.sum())
    # the weight schedule: lists as guided_inference builds them (:336-366), looked up by the reference CLASS (:622-665)
    T = 38
    for kind in ("constant", "linear", "quadratic"):
        for fg_w, bg_w in ((1.5, 1.25), (1.0, 2.0)):
            fg, bg = fg_w * 30, bg_w * 30
            if kind == "constant":
                ff, bf = np.linspace(fg, fg, T), np.linspace(bg, bg, T)
            elif kind == "linear":
                ff, bf = np.linspace(fg, 0.0, T), np.linspace(bg, 0.0, T)
            else:
                ff, bf = np.linspace(np.sqrt(fg), 0.0, T) ** 2, np.linspace(np.sqrt(bg), 0.0, T) ** 2
            pattern = {0: ([0.0, 0.0, 7.5], [0.0, 0.0, 1.5]), 1: ([0.0, 5.0, 0.0], [0.0, 1.5, 0.0]), 2: ([0.0, 5.0, 7.5], [0.0, 1.5, 1.5])}
            den = [(t, (np.array(pattern[t % 3][0]) * ff[t]).tolist(), (np.array(pattern[t % 3][1]) * bf[t]).tolist()) for t in range(T)]
            den.append((T, [0.0] * 3, [0.0] * 3))
            opt = [(0, [2.5] * 3, [1.25] * 3), (1, [1.25] * 3, [2.5] * 3), (2, [1.25] * 3, [1.25] * 3), (3, [2.5] * 3, [2.5] * 3)]
            sched = ref.gsd.StepGuidanceWeightSchedule(denoising_steps=den, optimization_steps=opt)
            table = np.array([[sched(t, it) for it in range(6)] for t in range(52)], dtype=np.float64)      # (52, 6, 2, 3)
            out[f"schedule/{kind}_{fg_w}_{bg_w}"] = table
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pins.npz"), **out)
    print(f"[pins] {len(out)} arrays")


def main():
    ref = load_reference()
    which = sys.argv[1:] or ["photogen", "pins"]
    if "photogen" in which:
        make_photogen(ref)
    if "pins" in which:
        make_pins(ref)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
