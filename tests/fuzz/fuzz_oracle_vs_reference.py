"""CPU only, where /root/reference exists: the oracle against the REAL reference on random inputs (beyond the committed
golden vectors).  transform_depth_pc (512^2, the reference is hard-wired to that size), process_correspondences, losses.
python tests/fuzz/fuzz_oracle_vs_reference.py [n_cases] [seed]"""
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dh_oracle as O                                        # noqa: E402
from oracle.ref_loader import load_reference                             # noqa: E402

warnings.filterwarnings("ignore")


def run(n_cases: int, seed: int) -> int:
    ref = load_reference()
    rng = np.random.default_rng(seed)
    K = ref.get_depth_intrinsics()
    bad = 0
    for case in range(n_cases):
        # ---- geometry ----
        kind = "smooth" if rng.random() < 0.4 else "disc"
        sseed = int(rng.integers(0, 10_000))
        scene = dict(S=512, seed=sseed, kind="smooth") if kind == "smooth" else dict(
            S=512, seed=sseed, cx=float(rng.uniform(100, 400)), cy=float(rng.uniform(100, 400)), radius=float(rng.uniform(15, 200)),
            quantize=float(rng.choice([0.0, 0.0, 0.1])) or None)
        depth, bg, mask = O.synthetic_scene(**scene)
        axis = [(0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -3.0, 0.0)][rng.integers(0, 4)]
        angle = float(rng.uniform(-180, 180))
        t = tuple(float(v) for v in rng.normal(size=3) * (2.0 if rng.random() < 0.25 else 0.3))
        norm = bool(rng.random() < 0.3)
        disp, corr = ref.depth_transform.transform_depth_pc(
            torch.from_numpy(depth)[None, None], torch.from_numpy(bg)[None, None], torch.from_numpy(mask)[None, None], K,
            rot_angle=angle, rot_axis=torch.tensor(axis, dtype=torch.float32), translation=torch.tensor(t, dtype=torch.float32),
            use_input_depth_normalization=norm)
        t32 = tuple(float(np.float32(v)) for v in t)
        o = O.transform_depth_pc(depth, bg, mask, K.numpy(), angle, axis, t32, use_input_depth_normalization=norm)
        ok = np.array_equal(corr.numpy(), o["correspondences"]) and np.array_equal(disp[0, 0].numpy(), o["disparity"], equal_nan=True)
        if not ok:
            bad += 1
            print(f"MISMATCH geometry case {case}: {scene} axis={axis} angle={angle} t={t} norm={norm}", flush=True)
        # ---- process_correspondences + losses on this edit's correspondences ----
        c = corr.numpy()
        if len(c) > 4000:
            c = c[np.sort(rng.choice(len(c), 4000, replace=False))]
        er = int(rng.choice([0, 0, 2, 9]))
        if len(c) == 1:          # the reference itself raises on exactly one correspondence (.squeeze() -> 0-d tensors, :494-510)
            continue
        pc_ref = ref.process_correspondences(torch.from_numpy(c), 512, er)
        pc = O.process_correspondences(c, 512, er)
        if not all(np.array_equal(np.asarray(pc_ref[k]), pc[k]) for k in pc):
            bad += 1
            print(f"MISMATCH process_correspondences case {case}", flush=True)
        if len(c) == 0:
            continue
        C, h = int(rng.integers(1, 6)), int(rng.choice([16, 32, 64]))
        patch = int(rng.choice([1, 1, 2, 3]))
        cur = torch.from_numpy(rng.normal(size=(C, h, h))).requires_grad_(True)      # fp64: no sign ambiguity to speak of
        orig = torch.from_numpy(rng.normal(size=(C, h, h)))
        for name, fn_ref, fn_o in (
                ("fg", lambda: ref.losses.compute_foreground_loss(cur, orig, pc_ref, patch, (64, 64)),
                 lambda: O.foreground_loss(cur.detach().numpy(), orig.numpy(), pc, patch=patch)),
                ("bg_global", lambda: ref.losses.compute_background_loss(cur, orig, pc_ref, patch, (64, 64)),
                 lambda: O.background_loss(cur.detach().numpy(), orig.numpy(), pc)),
                ("bg_local", lambda: ref.losses.compute_background_loss(cur, orig, pc_ref, patch, (64, 64), 'local_avg'),
                 lambda: O.background_loss(cur.detach().numpy(), orig.numpy(), pc, loss_type="local_avg", patch=patch))):
            l = fn_ref()
            g = torch.autograd.grad(l, cur)[0].numpy()
            v, go = fn_o()
            if not (abs(float(l) - v) <= 1e-12 * max(abs(v), 1) and np.abs(g - go).max() <= 1e-12):
                bad += 1
                print(f"MISMATCH loss {name} case {case}: C={C} h={h} patch={patch}: {float(l)} vs {v}, grad {np.abs(g - go).max()}", flush=True)
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    t0 = time.time()
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(f"{n} cases, {bad} mismatching, {time.time() - t0:.0f} s")
