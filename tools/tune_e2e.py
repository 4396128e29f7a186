"""End-to-end (pinned host -> GPU -> pinned host) sweep of EditWarpPipeline chunk size / stream count, next to the raw
PCIe copy rates of the box.  python tools/tune_e2e.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload, LEVELS, S, EDITS_PER_GPU                # noqa: E402
from diffusionhandles_b200.batch import EditWarpPipeline                  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n = EDITS_PER_GPU

# raw PCIe: 1 GiB pinned each way, alone and both directions at once
h_in = torch.empty(1 << 28, dtype=torch.float32).pin_memory()
h_out = torch.empty(1 << 28, dtype=torch.float32).pin_memory()
d_a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
d_b = torch.empty(1 << 28, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


gb = h_in.numel() * 4 / 1e9
t_h2d, t_d2h = timed(h2d), timed(d2h)
t_both = timed(lambda: (h2d(), d2h()))
print(json.dumps({"pcie": {"h2d_GBs": gb / t_h2d, "d2h_GBs": gb / t_d2h, "both_h2d_GBs": gb / t_both, "both_d2h_GBs": gb / t_both}}), flush=True)
del h_in, h_out, d_a, d_b

wl = build_workload(dev, n, full_map=True)
levels_h = [torch.randn((n, c, s, s), dtype=torch.float32).pin_memory() for c, s in LEVELS]
outs_h = [torch.empty((n, c, s, s), dtype=torch.float32).pin_memory() for c, s in LEVELS]
n_corr_h = torch.empty(n, dtype=torch.int32).pin_memory()
for chunk, streams in [(16, 3), (8, 3), (8, 4), (16, 4), (4, 4), (4, 6), (8, 6), (32, 3)]:
    pipe = EditWarpPipeline(dev, S, LEVELS, chunk=chunk, n_streams=streams, full_winner_map=True)
    t = timed(lambda: pipe.run_host(wl["scene_h"][0], wl["scene_h"][1], wl["scene_h"][2], wl["K"], wl["rigids"], levels_h, outs_h, n_corr_h,
                                    scene_index=wl["scene_index"]))
    bytes_in, bytes_out = pipe.h2d_bytes_per_edit(edits_per_scene=16) * n, pipe.d2h_bytes_per_edit() * n
    print(json.dumps({"chunk": chunk, "streams": streams, "warps_per_s": n / t, "h2d_GBs": bytes_in / t / 1e9,
                      "d2h_GBs": bytes_out / t / 1e9}), flush=True)
    del pipe
