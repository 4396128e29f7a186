"""Runs the batched geometry half of the path (K1 -> K2 -> masks -> correspondences -> dense maps) a few times, for
`ncu --metrics gpu__time_duration.sum` launch lists:  ncu ... python tools/profile_geometry.py [batch] [S]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import edit_recipe, LEVELS                                     # noqa: E402
from diffusionhandles_b200 import warp                                   # noqa: E402
from diffusionhandles_b200.engine import EditEngine, make_rigid          # noqa: E402
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser   # noqa: E402
from diffusionhandles_b200.synthetic import synthetic_scene              # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = 512
dev = torch.device("cuda", 0)
scenes, edits = edit_recipe(B)
K = GuidedStableDiffuser.get_depth_intrinsics()
sc = [synthetic_scene(**s) for s in scenes]
depth = torch.stack([torch.from_numpy(sc[si][0]) for si, *_ in edits]).to(dev)
bg = torch.stack([torch.from_numpy(sc[si][1]) for si, *_ in edits]).to(dev)
mask = torch.stack([torch.from_numpy(sc[si][2]) for si, *_ in edits]).to(dev)
rigids = [make_rigid(a, list(ax), list(t)) for _, a, ax, t in edits]
eng = EditEngine(dev, B, S, S)
for it in range(3):
    res = eng.run(depth, bg, mask, K, rigids, poisson=False, sync_counts=False)
    maps = warp.dense_source_maps(res.corr, res.n_corr, S, [s for _, s in LEVELS], res.winner_src)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for it in range(20):
    res = eng.run(depth, bg, mask, K, rigids, poisson=False, sync_counts=False)
    maps = warp.dense_source_maps(res.corr, res.n_corr, S, [s for _, s in LEVELS], res.winner_src)
ev1.record()
torch.cuda.synchronize()
print(f"batch {B} S {S}: {ev0.elapsed_time(ev1) / 20:.3f} ms per batch, {B * 20 / ev0.elapsed_time(ev1) * 1e3:.0f} edits/s")
