"""One config-1 edit with the Poisson hole fill, for ncu captures of poisson_cg_kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionhandles_b200.engine import EditEngine, make_rigid
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
from diffusionhandles_b200.synthetic import synthetic_scene
dev = torch.device("cuda:0")
K = GuidedStableDiffuser.get_depth_intrinsics()
depth, bg, mask = synthetic_scene(512, 0)
eng = EditEngine(dev, 1, 512, 512)
td, tb, tm = (torch.from_numpy(a).to(dev)[None].contiguous() for a in (depth, bg, mask))
for _ in range(3):
    res = eng.run(td, tb, tm, K, [make_rigid(30.0, [0.0, 1.0, 0.0], [0.3, 0.0, 0.2])], poisson=True)
torch.cuda.synchronize()
print("iters", int(eng.poisson_iters[0]), "unknowns", int((eng.unpack_bits(res.cleaned_bits)[0] != res.target_mask[0]).sum()))
