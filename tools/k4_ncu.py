"""Minimal K4 driver for ncu captures: config-3 shapes, a few fused loss launches.
ncu --set full --clock-control none --import-source on -k regex:loss_ --launch-skip 4 --launch-count 2 -o gpurun_out/k4 python tools/k4_ncu.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionhandles_b200 import losses
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
dev = torch.device("cuda:0")
gp = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/pc_transform.npz"))
corr = gp["cfg1/corr"].astype(np.int64)
pc = GuidedStableDiffuser().process_correspondences(torch.from_numpy(corr), 512, 0)
shapes = [(1280, 32), (640, 64), (320, 64)]
g = torch.Generator(device=dev).manual_seed(3)
curs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
origs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
plan = losses._plan_for(pc, 64, dev)
for _ in range(4):
    losses._launch(curs, origs, [True] * 3, [1.0] * 3, [1.0] * 3, plan, 1, 1)
torch.cuda.synchronize()
