"""Device time of a small resident sweep (the per-rank body of the 8-GPU strong-scaling run: 32 edits) for different numbers
of parallel graph branches.  python tools/sweep_branches.py [n_edits]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from diffusionhandles_b200.batch import DeviceSweep, shard_edits
from diffusionhandles_b200.synthetic import synthetic_scene
from diffusionhandles_b200.engine import make_rigid
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
world = 256 // n
scenes, edits = bench.edit_recipe(256)
mine = shard_edits(256, 0, world)
K = GuidedStableDiffuser.get_depth_intrinsics()
sc = {}
for e in mine:
    si = edits[e][0]
    if si not in sc:
        sc[si] = [torch.from_numpy(a).to(dev) for a in synthetic_scene(**scenes[si])]
depth, bg, mask = (torch.stack([sc[edits[e][0]][k] for e in mine]).contiguous() for k in range(3))
rigids = [make_rigid(edits[e][1], list(edits[e][2]), list(edits[e][3])) for e in mine]
levels = [torch.randn((n, c, s, s), device=dev) for c, s in bench.LEVELS]
ref = None
for chunk, br in ((n, 1), (n, 2), (n, 4), (n, 8), (n // 2, 1), (n // 2, 2)):
    sw = DeviceSweep(dev, 512, bench.LEVELS, depth, bg, mask, K, rigids, levels, chunk=chunk, use_graph=True, branches=br)
    sw.capture()
    for _ in range(3):
        sw.run()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        a.record(); sw.run(); b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    if ref is None:
        ref = sw.rec.clone()
    assert torch.equal(ref, sw.rec)
    print(json.dumps({"edits": n, "chunk": chunk, "branches": br, "ms": ms, "edits_per_s": n / ms * 1e3}), flush=True)
    del sw
