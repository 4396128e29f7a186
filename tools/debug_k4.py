import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dh_oracle as O
from diffusionhandles_b200 import losses
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
dev = torch.device("cuda:0")
gp = np.load("tests/golden/pc_transform.npz")
corr = gp["cfg1/corr"].astype(np.int64)
pc = GuidedStableDiffuser().process_correspondences(torch.from_numpy(corr), 512, 0)
pc_np = O.process_correspondences(corr, 512, 0)
shapes = [(1280, 32), (640, 64), (320, 64)]
g3, g4 = torch.Generator().manual_seed(3), torch.Generator().manual_seed(4)
curs = [torch.randn((c, s, s), generator=g3) for c, s in shapes]
origs = [torch.randn((c, s, s), generator=g4) for c, s in shapes]
fgw, bgw = [3.0, 5.0, 7.5], [2.0, 1.5, 1.5]
for lt in ("global_avg", "local_avg"):
    dc = [c.to(dev).requires_grad_(True) for c in curs]
    total, parts = losses.guidance_loss(dc, [o.to(dev) for o in origs], pc, fgw, bgw, bg_loss_type=lt)
    grads = torch.autograd.grad(total, dc)
    for l, (c, o_) in enumerate(zip(curs, origs)):
        vf, gf = O.foreground_loss(c.numpy(), o_.numpy(), pc_np)
        vb, gb = O.background_loss(c.numpy(), o_.numpy(), pc_np, loss_type=lt)
        ref = fgw[l] * gf + bgw[l] * gb
        a = grads[l].cpu().numpy()
        d = np.abs(a - ref)
        print(lt, "layer", l, "fg", parts[2*l].item(), vf, "bg", parts[2*l+1].item(), vb, "grad maxdiff", d.max(), "ref max", np.abs(ref).max(),
              "argmax", np.unravel_index(d.argmax(), d.shape), "n_bad", int((d > 1e-5*np.abs(ref).max()).sum()))
