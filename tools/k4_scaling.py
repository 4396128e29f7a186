"""K4 kernel time versus channel count (fixed per-launch overhead vs per-channel cost), CUDA-graph timed."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionhandles_b200 import losses
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
dev = torch.device("cuda:0")
gp = np.load("tests/golden/pc_transform.npz")
pc = GuidedStableDiffuser().process_correspondences(torch.from_numpy(gp["cfg1/corr"].astype(np.int64)), 512, 0)
plan = losses._plan_for(pc, 64, dev)


def timed(shapes):
    curs = [torch.randn((c, s, s), device=dev) for c, s in shapes]
    origs = [torch.randn((c, s, s), device=dev) for c, s in shapes]
    L = len(shapes)
    fn = lambda: losses._launch(curs, origs, [True] * L, [1.0] * L, [1.0] * L, plan, 1, 1)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) / 20 * 1e3


for name, shapes in [("flat   C=4", [(4, 64)]), ("flat   C=148", [(148, 64)]), ("flat   C=444", [(444, 64)]), ("flat   C=960", [(960, 64)]),
                     ("flat   C=1920", [(1920, 64)]), ("resize C=4", [(4, 32)]), ("resize C=444", [(444, 32)]), ("resize C=1280", [(1280, 32)]),
                     ("resize C=2560", [(2560, 32)]), ("all three", [(1280, 32), (640, 64), (320, 64)])]:
    print(f"{name:16s} {timed(shapes):8.1f} us")
