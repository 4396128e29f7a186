"""Sweep of the K3 launch tunables on the GPU box (stages x CTAs/SM x segment size).  Prints one line per point."""
import itertools
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionhandles_b200 import warp  # noqa: E402

LEVELS = [(320, 64), (640, 32), (1280, 16), (1280, 8)]
B = 256
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
levels = [torch.randn((B, c, s, s), generator=gen, device=dev) for c, s in LEVELS]
outs = [torch.empty_like(l) for l in levels]
# rotation-like maps: shifted / mirrored identity with a few holes
maps = []
for c, s in LEVELS:
    q = torch.arange(s * s, device=dev)
    y, x = q // s, q % s
    m = torch.stack([(y * s + (x * (0.6 + 0.4 * (e % 16) / 15.0)).long().clamp(max=s - 1)) for e in range(B)]).to(torch.int32)
    maps.append(m.contiguous())
algo = sum(2 * c * s * s * 4 + s * s * 4 for c, s in LEVELS) * B


def run(n=12):
    for _ in range(3):
        warp.warp_stacks(levels, maps, outs)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); warp.warp_stacks(levels, maps, outs); b.record()
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(t)), float(np.min(t))


res = []
for stages, ctas, seg in itertools.product([2, 3, 4, 6, 8], [1, 2, 3, 4, 6], [5, 10, 20, 40]):
    if stages * 16 * ctas > 220:
        continue
    os.environ["DH_WARP_STAGES"], os.environ["DH_WARP_CTAS_PER_SM"], os.environ["DH_WARP_SEG_CHUNKS"] = str(stages), str(ctas), str(seg)
    med, mn = run()
    res.append(dict(stages=stages, ctas=ctas, seg=seg, med_ms=med, min_ms=mn, gbs=algo / med / 1e6))
    print(json.dumps(res[-1]), flush=True)
best = max(res, key=lambda r: r["gbs"])
print("BEST", json.dumps(best))
