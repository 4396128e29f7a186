"""Per-phase cycles of thread 0 of every group of the fused loss kernel, for a list of layer shapes - e.g. ONE resized item alone
(4x32) against the full layer (1280x32): what an item costs in isolation and under load (profiles/r02_k4_summary.md).
Needs a library built with the timers: `DH_LOSS_PHASE_TIMERS=1 python -m diffusionhandles_b200.build --force` (rebuild without
the variable afterwards), or a side build pointed at by DH_B200_LIB.
usage: python tools/k4_phases.py 4x32 1280x32 4x64 960x64 all"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
dbg = torch.zeros(12 * 1024, dtype=torch.int64, device=dev)
os.environ["DH_LOSS_DEBUG_BUF"] = str(dbg.data_ptr())
from diffusionhandles_b200 import losses
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
gp = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/pc_transform.npz"))
pc = GuidedStableDiffuser().process_correspondences(torch.from_numpy(gp["cfg1/corr"].astype(np.int64)), 512, 0)
plan = losses._plan_for(pc, 64, dev)
names = ["wait_tma", "flat_item", "small_upsample+bg", "small_rows", "small_reduce", "small_stage_next", "prologue", "small_grad_gather"]
for arg in sys.argv[1:]:
    shapes = [(1280, 32), (640, 64), (320, 64)] if arg == "all" else [tuple(int(v) for v in s.split("x")) for s in arg.split(",")]
    g = torch.Generator(device=dev).manual_seed(3)
    curs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
    origs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
    L = len(shapes)
    for _ in range(3):
        dbg.zero_()
        losses._launch(curs, origs, [True] * L, [1.0] * L, [1.0] * L, plan, 1, 1)
    torch.cuda.synchronize()
    raw = dbg.cpu().numpy()
    d = raw[:4096].reshape(-1, 4)
    n = int((d[:, 1] > 0).sum())
    d = d[:n]
    ph = raw[4096:4096 + 8 * n].reshape(-1, 8)
    act = d[:, 2] > 0
    print(f"== {arg}: groups {n}, with items {int(act.sum())}, items/group max {int(d[:, 2].max())}; kernel span {(d[:, 1].max() - d[:, 0].min()) / 1e3:.1f} us")
    if ph.any() and act.any():
        a = ph[act]
        per_item = a / np.maximum(d[act, 2:3], 1)
        for i, nm in enumerate(names):
            print(f"  {nm:>18}: median/group {np.median(a[:, i]):9.0f} cyc   per item {np.median(per_item[:, i]):9.0f}   share {100 * a[:, i].sum() / a.sum():5.1f}%")
        print(f"  {'total':>18}: median/group {np.median(a.sum(1)):9.0f} cyc")
