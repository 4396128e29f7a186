"""Randomised parity sweep of the guidance losses (K4, both kernels) against the fp64 oracle: random index lists with
duplicates, random layer shapes (including non-square and non-power-of-two maps), both background types, patch sizes 1..5.
python tests/fuzz/fuzz_losses.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dh_oracle as O                                        # noqa: E402
from diffusionhandles_b200 import losses                                 # noqa: E402


def run(n_cases: int, seed: int, verbose: bool = True) -> int:
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    bad = 0
    for case in range(n_cases):
        n = int(rng.choice([1, 7, 300, 3000, 20000]))
        # correspondences at the 64-grid: sources and destinations in random boxes, duplicates guaranteed for large n
        bx0, by0 = rng.integers(0, 40, 2)
        bw, bh = rng.integers(1, 64 - bx0 + 1), rng.integers(1, 64 - by0 + 1)
        ox, oy = rng.integers(bx0, bx0 + bw, n), rng.integers(by0, by0 + bh, n)
        tx, ty = rng.integers(0, 64, n), rng.integers(0, 64, n)
        bgm = rng.random((64, 64)) < rng.uniform(0.05, 0.9)
        bgo, bgt = rng.random((64, 64)) < 0.5, rng.random((64, 64)) < 0.5
        if not bgm.any(): bgm[3, 5] = True
        if not bgo.any(): bgo[1, 1] = True
        if not bgt.any(): bgt[2, 2] = True
        pc = dict(original_x=ox, original_y=oy, transformed_x=tx, transformed_y=ty,
                  background_y=np.nonzero(bgm)[0], background_x=np.nonzero(bgm)[1],
                  background_y_orig=np.nonzero(bgo)[0], background_x_orig=np.nonzero(bgo)[1],
                  background_y_trans=np.nonzero(bgt)[0], background_x_trans=np.nonzero(bgt)[1])
        C = int(rng.integers(1, 24))
        h, w = [(64, 64), (32, 32), (16, 16), (8, 8), (48, 48), (32, 48), (64, 32), (24, 24), (4, 4), (12, 20)][rng.integers(0, 10)]
        patch = int(rng.choice([1, 1, 1, 2, 3, 5]))
        lt = ["global_avg", "local_avg"][rng.integers(0, 2)]
        fgw, bgw = float(rng.uniform(0.1, 5)), float(rng.uniform(0.1, 5))
        cur = rng.normal(size=(C, h, w)).astype(np.float32)
        orig = rng.normal(size=(C, h, w)).astype(np.float32)
        vf, gf = O.foreground_loss(cur, orig, pc, patch=patch)
        vb, gb = O.background_loss(cur, orig, pc, loss_type=lt, patch=patch)
        ref_v, ref_g = fgw * vf + bgw * vb, fgw * gf + bgw * gb
        amb = O.loss_sign_ambiguity(cur, orig, pc, bg_loss_type=lt, patch=patch)
        tc = torch.from_numpy(cur).to(dev).requires_grad_(True)
        total, parts = losses.guidance_loss([tc], [torch.from_numpy(orig).to(dev)], pc, [fgw], [bgw], bg_loss_type=lt, patch_size=patch)
        g = torch.autograd.grad(total, tc)[0].cpu().numpy()
        ok_v = abs(total.item() - ref_v) <= 1e-5 * abs(ref_v)
        scale = max(np.abs(ref_g).max(), 1e-30)
        err = np.where(amb, 0.0, np.abs(g - ref_g)).max()
        ok_g = err <= 1e-5 * scale
        if not (ok_v and ok_g):
            bad += 1
            if verbose:
                print(f"MISMATCH case {case}: n={n} C={C} h={h} w={w} patch={patch} {lt}: value {total.item()} vs {ref_v}, "
                      f"grad err {err / scale:.2e} (ambiguous {amb.mean():.1e})", flush=True)
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    t0 = time.time()
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(f"{n} cases, {bad} mismatching, {time.time() - t0:.0f} s")
