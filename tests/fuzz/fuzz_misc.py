"""Randomised parity sweeps of the remaining integer / copy pieces against the oracle (bit-exact):
  p2d    points_to_depth on arbitrary fp64 point sets (duplicates, exact z ties, points behind the camera, off-screen, NaN z)
  pcorr  process_correspondences (random lists, image sizes, erosion, out-of-bounds entries)
  warp   gather_list / warp_stacks on random shapes (TMA fast path and the generic path) against torch indexing
python tests/fuzz/fuzz_misc.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dh_oracle as O                                        # noqa: E402
from diffusionhandles_b200 import depth_transform as dt, warp           # noqa: E402
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser   # noqa: E402

dev = torch.device("cuda:0")
K = GuidedStableDiffuser.get_depth_intrinsics()
K_NP = K.numpy()
KEYS = ['original_x', 'original_y', 'transformed_x', 'transformed_y', 'background_x', 'background_y',
        'background_x_orig', 'background_y_orig', 'background_x_trans', 'background_y_trans']


def fuzz_p2d(rng) -> bool:
    H, W = [(32, 32), (64, 48), (17, 33), (128, 128), (256, 200)][rng.integers(0, 5)]
    n = int(rng.choice([1, 50, 5000, 60000]))
    xy = rng.uniform(-1.6, 1.6, (n, 2))
    z = rng.uniform(0.5, 6.0, n)
    if rng.random() < 0.5:
        z = np.round(z * 4) / 4                       # exact z ties
    if rng.random() < 0.5:
        xy = np.round(xy * 20) / 20                   # many points per pixel
    behind = rng.random(n) < 0.05
    z[behind] *= -1.0                                 # behind the camera: a negative z beats every positive one
    pts = np.concatenate([xy * z[:, None] / K_NP[0, 0], z[:, None]], axis=1)
    pm = (rng.random(n) < 0.4).astype(np.uint8)
    dm, mk, tx, ty, vis = dt.points_to_depth(torch.from_numpy(pts).to(dev), K, (H, W), point_mask=torch.from_numpy(pm).to(dev))
    odm, omk, otx, oty, ovis, _ = O.points_to_depth(pts, K_NP, (H, W), pm)
    return (np.array_equal(dm[0, 0].cpu().numpy(), odm) and np.array_equal(mk, omk) and np.array_equal(vis, ovis)
            and np.array_equal(tx, otx) and np.array_equal(ty, oty))


def fuzz_pcorr(rng) -> bool:
    res = int(rng.choice([64, 128, 256, 512, 1024]))
    n = int(rng.choice([0, 1, 40, 3000, 30000]))
    corr = rng.integers(0, res, (n, 4))
    if n and rng.random() < 0.5:                      # concentrated lists: many duplicates per cell
        corr[:, 2:] = corr[:, 2:] // 4 + res // 3
    if n and rng.random() < 0.3:                      # a few out-of-bounds rows (filtered, guided_stable_diffuser.py:509-514)
        k = min(n, 5)
        corr[:k, rng.integers(2, 4)] = rng.choice([-1, res, res + 7])   # destination columns only: sources are always pixels
    er = int(rng.choice([0, 0, 1, 3, 8, 20]))
    pc = GuidedStableDiffuser().process_correspondences(torch.from_numpy(corr.astype(np.int64)), res, er)
    ref = O.process_correspondences(corr.astype(np.int64), res, er)
    return all(np.array_equal(pc[k], ref[k]) and pc[k].dtype == np.int64 for k in KEYS)


def fuzz_warp(rng) -> bool:
    B = int(rng.integers(1, 4))
    shapes = [[(320, 64), (640, 32), (1280, 16), (1280, 8)], [(5, 64), (3, 32)], [(7, 24), (2, 10), (9, 3)], [(33, 16), (1, 8)]][rng.integers(0, 4)]
    levels = [torch.randn((B, c, s, s), device=dev) for c, s in shapes]
    maps = []
    for c, s in shapes:
        m = torch.randint(-1, s * s, (B, s * s), device=dev, dtype=torch.int32)
        if rng.random() < 0.3:
            m[:] = -1
        maps.append(m)
    outs = warp.warp_stacks(levels, maps)
    ok = True
    for a, m, o in zip(levels, maps, outs):
        idx = m.long().clamp(min=0)
        ref = torch.gather(a.flatten(2), 2, idx[:, None, :].expand(-1, a.shape[1], -1)) * (m >= 0)[:, None, :]
        ok &= bool(torch.equal(o.flatten(2), ref))
    C, s = int(rng.integers(1, 40)), int(rng.choice([8, 16, 31, 64]))
    A = torch.randn((C, s, s), device=dev)
    n = int(rng.choice([0, 1, 17, 5000]))
    y, x = rng.integers(0, s, n), rng.integers(0, s, n)
    g = warp.gather_list(A, y, x)
    ok &= bool(torch.equal(g, A[:, torch.from_numpy(y).to(dev), torch.from_numpy(x).to(dev)]))
    return ok


def run(n_cases: int, seed: int, verbose: bool = True) -> int:
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        for name, fn in (("p2d", fuzz_p2d), ("pcorr", fuzz_pcorr), ("warp", fuzz_warp)):
            state = rng.bit_generator.state
            if not fn(rng):
                bad += 1
                if verbose:
                    print(f"MISMATCH {name} case {case} (seed {seed})", flush=True)
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    t0 = time.time()
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(f"{n} cases x 3 sweeps, {bad} mismatching, {time.time() - t0:.0f} s")
