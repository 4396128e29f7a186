"""Per-group timeline of the fused loss kernel (start / end / items / SM) through the DH_LOSS_DEBUG_BUF developer hook, and -
with a library built by `DH_LOSS_PHASE_TIMERS=1 python -m diffusionhandles_b200.build --force` - the cycles thread 0 of every
group spends in each phase (profiles/r02_k4_summary.md).  Rebuild without the variable afterwards: the timers cost time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
dbg = torch.zeros(12 * 1024, dtype=torch.int64, device=dev)
os.environ["DH_LOSS_DEBUG_BUF"] = str(dbg.data_ptr())
from diffusionhandles_b200 import losses
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
gp = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/pc_transform.npz"))
corr = gp["cfg1/corr"].astype(np.int64)
pc = GuidedStableDiffuser().process_correspondences(torch.from_numpy(corr), 512, 0)
shapes = [(1280, 32), (640, 64), (320, 64)]
g = torch.Generator(device=dev).manual_seed(3)
curs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
origs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
plan = losses._plan_for(pc, 64, dev)
for _ in range(5):
    losses._launch(curs, origs, [True] * 3, [1.0] * 3, [1.0] * 3, plan, 1, 1)
torch.cuda.synchronize()
raw = dbg.cpu().numpy()
d = raw[:4096].reshape(-1, 4)
n_cta = int((d[:, 1] > 0).sum())
d = d[:n_cta]
ph = raw[4096:4096 + 8 * n_cta].reshape(-1, 8)
t0 = d[:, 0].min()
start, end, items, sm = (d[:, 0] - t0) / 1e3, (d[:, 1] - t0) / 1e3, d[:, 2], d[:, 3]
print(f"CTAs {len(d)}  start us: min {start.min():.1f} med {np.median(start):.1f} max {start.max():.1f}")
print(f"end us: min {end.min():.1f} med {np.median(end):.1f} max {end.max():.1f}")
print(f"items per CTA: min {items.min()} med {np.median(items)} max {items.max()} sum {items.sum()}")
dur = end - start
print(f"duration us: min {dur.min():.1f} med {np.median(dur):.1f} max {dur.max():.1f};  us per item: {np.median(dur / np.maximum(items, 1)):.2f}")
order = np.argsort(end)
for i in list(order[:3]) + list(order[-3:]):
    print(f"  cta {i}: sm {sm[i]} start {start[i]:.1f} end {end[i]:.1f} items {items[i]}")
per_sm = {}
for s_, e_ in zip(sm, end):
    per_sm[s_] = max(per_sm.get(s_, 0), e_)
v = np.array(list(per_sm.values()))
print(f"SMs {len(v)}: last CTA end per SM: min {v.min():.1f} med {np.median(v):.1f} max {v.max():.1f}")

if ph.any():
    names = ["loop_top+wait_tma", "flat_item", "small_upsample+bg", "small_rows", "small_reduce", "small_stage_next", "prologue", "small_grad_gather"]
    tot = ph.sum(0)
    print("thread-0 cycles by phase (sum over CTAs, share):")
    for n, v in zip(names, tot):
        print(f"  {n:>14}: {v / 1e6:8.2f} M  {100 * v / tot.sum():5.1f}%")
    print(f"  per CTA total cycles med {np.median(ph.sum(1)):.0f}")
