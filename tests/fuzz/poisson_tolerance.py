"""Sweep of the hole-fill stopping tolerance against the SuperLU result (test tooling: uses the oracle)."""
import os, sys, numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
from oracle import dh_oracle as O
import fuzz_more as F
from diffusionhandles_b200 import engine as E
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
dev = torch.device("cuda:0"); K = GuidedStableDiffuser.get_depth_intrinsics(); K_NP = K.numpy()
rng = np.random.default_rng(3)
cases = []
for _ in range(40):
    S = int(rng.choice([96, 128, 192, 256]))
    d, b, m, a, ax, t = F.random_edit(rng, S)
    if m.any(): cases.append((S, d, b, m, a, ax, t))
for tol in (1e-13, 1e-11, 1e-10, 1e-9):
    errs, its = [], []
    for S, d, b, m, a, ax, t in cases:
        o = O.transform_depth_pc(d, b, m, K_NP, a, ax, tuple(float(np.float32(v)) for v in t), poisson=True)
        e = E.EditEngine(dev, 1, S, S)  # private engine: the cached one keeps its tolerance
        td, tb, tm = (torch.from_numpy(x).to(dev)[None].contiguous() for x in (d, b, m))
        e.poisson_rel_tol = tol
        res = e.run(td, tb, tm, K, [E.make_rigid(a, list(ax), list(t))], poisson=True)
        errs.append(np.abs(res.disparity[0].cpu().numpy() - o["disparity"]).max()); its.append(int(e.poisson_iters[0]))
    print(f"tol {tol:g}: max err {max(errs):.2e}, mean iters {np.mean(its):.1f}")
