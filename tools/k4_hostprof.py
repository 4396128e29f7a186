"""Host-side profile of one guidance-loss evaluation through the public API (autograd included)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionhandles_b200 import losses
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
dev = torch.device("cuda:0")
gp = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/pc_transform.npz"))
pc = GuidedStableDiffuser().process_correspondences(torch.from_numpy(gp["cfg1/corr"].astype(np.int64)), 512, 0)
shapes = [(1280, 32), (640, 64), (320, 64)]
g = torch.Generator(device=dev).manual_seed(3)
curs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
origs = [torch.randn((c, s, s), generator=g, device=dev) for c, s in shapes]
fgw, bgw = [1.0, 2.0, 3.0], [1.0, 1.5, 2.0]


def evaluation():
    cs = [c.detach().requires_grad_(True) for c in curs]
    total, _ = losses.guidance_loss(cs, origs, pc, fgw, bgw)
    return torch.autograd.grad(total, cs)


for _ in range(20):
    evaluation()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    evaluation()
torch.cuda.synchronize()
print(f"wall per evaluation: {(time.perf_counter() - t0) / n * 1e6:.1f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    evaluation()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
