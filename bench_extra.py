#!/usr/bin/env python
"""Secondary measurements (BASELINE configs 1, 3, 4, 5) - one JSON object per line.  Not the driver's bench;
results are copied into profiles/ and quoted in DESIGN.md.

  config1  single 512^2 edit: latency of K1 -> K2 -> masks -> correspondences (+ Poisson) and of each stage
  config3  guidance loss fwd+bwd, recorded-stack shapes, 150 evaluations, vs 62,914,560 algorithmic bytes
  config4  edits/s of the geometry path for a 256-edit sweep (device-resident inputs)
  config5  1024^2 stress edit latency
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from diffusionhandles_b200 import losses, warp
from diffusionhandles_b200.engine import EditEngine, make_rigid
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser, make_guidance_weight_schedule
from diffusionhandles_b200.synthetic import synthetic_scene

dev = torch.device("cuda:0")
PEAK = 6535.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(t)), float(np.min(t))


def config1_and_5():
    K = GuidedStableDiffuser.get_depth_intrinsics()
    for name, S, scene, angle, t in (("config1", 512, dict(S=512, seed=0), 30.0, (0.3, 0.0, 0.2)),
                                     ("config5A", 1024, dict(S=1024, seed=0, cx=512.0, cy=560.0, radius=300.0), 90.0, (1.5, 0.0, 1.0)),
                                     ("config5C", 1024, dict(S=1024, seed=0, cx=512.0, cy=560.0, radius=300.0), 60.0, (-2.0, 0.0, -1.5))):
        depth, bg, mask = synthetic_scene(**scene)
        eng = EditEngine(dev, 1, S, S)
        td, tb, tm = (torch.from_numpy(a).to(dev)[None].contiguous() for a in (depth, bg, mask))
        rg = [make_rigid(angle, [0.0, 1.0, 0.0], list(t))]
        res = eng.run(td, tb, tm, K, rg, poisson=True)
        med_np, _ = timeit(lambda: eng.run(td, tb, tm, K, rg, poisson=False, sync_counts=False))
        med_p, _ = timeit(lambda: eng.run(td, tb, tm, K, rg, poisson=True, sync_counts=False))
        t0 = time.perf_counter()
        for _ in range(20):
            eng.run(td, tb, tm, K, rg, poisson=False, sync_counts=True)
        wall = (time.perf_counter() - t0) / 20 * 1e3
        extra = {}
        if name == "config1":
            # BASELINE config 1: "+ one 64x64x320 activation" warped through this edit (dense form through the winner map, and the
            # list form A[:, y_src, x_src] that the reference's losses gather)
            A = torch.randn((1, 320, 64, 64), generator=torch.Generator(device=dev).manual_seed(1), device=dev)
            out = torch.empty_like(A)
            maps = warp.dense_source_maps(res.corr, res.n_corr, S, [64], res.winner_src)
            extra["gpu_ms_warp_320x64x64_dense"], _ = timeit(lambda: warp.warp_stacks([A], maps, [out]))
            c = res.correspondences(0)
            ys, xs = (c[:, 1] // (S // 64)), (c[:, 0] // (S // 64))
            extra["gpu_ms_warp_320x64x64_list"], _ = timeit(lambda: warp.gather_list(A[0], ys, xs))
            extra["gpu_ms_maps"], _ = timeit(lambda: warp.dense_source_maps(res.corr, res.n_corr, S, [64], res.winner_src))
        print(json.dumps({"config": name, "S": S, "n_fg": int(res.n_fg_host[0]), "n_corr": int(res.n_corr_host[0]),
                          "gpu_ms_geometry": med_np, "gpu_ms_with_poisson": med_p, "wall_ms_with_count_readback": wall,
                          "poisson_iters": int(eng.poisson_iters[0].item()), **extra}), flush=True)
        del eng


def config3():
    K = GuidedStableDiffuser.get_depth_intrinsics()
    depth, bg, mask = synthetic_scene(512, 0)
    eng = EditEngine(dev, 1, 512, 512)
    res = eng.run(*(torch.from_numpy(a).to(dev)[None].contiguous() for a in (depth, bg, mask)), K,
                  [make_rigid(30.0, [0.0, 1.0, 0.0], [0.3, 0.0, 0.2])], poisson=False)
    pc = GuidedStableDiffuser().process_correspondences(res.correspondences(0), 512, 0)
    shapes = [(1280, 32), (640, 64), (320, 64)]
    T = 50
    g3, g4 = torch.Generator(device=dev).manual_seed(3), torch.Generator(device=dev).manual_seed(4)
    origs = [torch.randn((T, c, s, s), generator=g4, device=dev) for c, s in shapes]       # recorded stacks, 1.05 GB
    curs = [torch.randn((3, c, s, s), generator=g3, device=dev) for c, s in shapes]
    sched = make_guidance_weight_schedule(1.5, 1.25)
    algo = 3 * sum(c * s * s for c, s in shapes) * 4
    state = {"i": 0}

    def evaluation():
        i = state["i"]; state["i"] += 1
        t_idx, it = (i // 3) % T, i % 3
        fgw, bgw = sched(min(t_idx, 37), it)
        fgw = [w if w else 1.0 for w in fgw]; bgw = [w if w else 1.0 for w in bgw]      # all three layers active
        cs = [c[it].requires_grad_(True) for c in curs]
        total, _ = losses.guidance_loss(cs, [o[t_idx] for o in origs], pc, fgw, bgw)
        torch.autograd.grad(total, cs)
    med, mn = timeit(evaluation, n=150, warm=6)
    # kernel-only: the fused launch without autograd plumbing
    plan = losses._plan_for(pc, 64, dev)
    fixed = [c[0] for c in curs]
    o0 = [o[0] for o in origs]

    def kernel_only():
        losses._launch(fixed, o0, [True] * 3, [1.0] * 3, [1.0] * 3, plan, 1, 1)
    kernel_only()
    torch.cuda.synchronize()
    # GPU time without Python launch overhead: 20 evaluations captured in one CUDA graph
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        kernel_only()
        with torch.cuda.graph(gr, stream=side):
            for _ in range(20):
                kernel_only()
    torch.cuda.current_stream().wait_stream(side)
    gmed, gmn = timeit(gr.replay, n=20, warm=3)
    kmed, kmn = gmed / 20, gmn / 20
    pmed, _ = timeit(kernel_only, n=50, warm=5)
    print(json.dumps({"config": "config3", "evaluations": 150, "n_corr": int(res.n_corr_host[0]),
                      "ms_per_evaluation_autograd": med, "ms_per_evaluation_kernels": kmed, "ms_per_evaluation_python_launch": pmed,
                      "algorithmic_bytes": algo, "achieved_gbs_kernels": algo / kmed / 1e6, "frac_of_hbm_peak": algo / kmed / 1e6 / PEAK}), flush=True)


def config4():
    import bench
    scenes, edits = bench.edit_recipe(256)
    K = GuidedStableDiffuser.get_depth_intrinsics()
    sc = [synthetic_scene(**s) for s in scenes]
    for chunk in (16, 64, 256):
        eng = EditEngine(dev, chunk, 512, 512)
        d = torch.stack([torch.from_numpy(sc[e[0]][0]) for e in edits]).to(dev)
        b = torch.stack([torch.from_numpy(sc[e[0]][1]) for e in edits]).to(dev)
        m = torch.stack([torch.from_numpy(sc[e[0]][2]) for e in edits]).to(dev)
        rg = [make_rigid(e[1], list(e[2]), list(e[3])) for e in edits]

        def sweep():
            for e0 in range(0, 256, chunk):
                r = eng.run(d[e0:e0 + chunk], b[e0:e0 + chunk], m[e0:e0 + chunk], K, rg[e0:e0 + chunk], poisson=False, sync_counts=False)
                warp.dense_source_maps(r.corr, r.n_corr, 512, [64, 32, 16, 8], r.winner_src)
        med, mn = timeit(sweep, n=5, warm=2)
        print(json.dumps({"config": "config4_geometry_sweep", "chunk": chunk, "edits": 256, "ms_per_sweep": med, "edits_per_s": 256 / med * 1e3}), flush=True)
        if chunk == 64:
            # the whole device-resident pipeline: geometry -> dense maps -> K3 warp of every edit's own config-2 stack
            levels = [torch.randn((256, c, s_, s_), device=dev) for c, s_ in bench.LEVELS]
            outs = [torch.empty_like(l) for l in levels]

            def full():
                for e0 in range(0, 256, chunk):
                    r = eng.run(d[e0:e0 + chunk], b[e0:e0 + chunk], m[e0:e0 + chunk], K, rg[e0:e0 + chunk], poisson=False, sync_counts=False)
                    maps = warp.dense_source_maps(r.corr, r.n_corr, 512, [s_ for _, s_ in bench.LEVELS], r.winner_src)
                    warp.warp_stacks([l[e0:e0 + chunk] for l in levels], maps, [o[e0:e0 + chunk] for o in outs])
            med, mn = timeit(full, n=5, warm=2)
            print(json.dumps({"config": "config4_full_pipeline_device_resident", "chunk": chunk, "edits": 256, "ms_per_sweep": med,
                              "edits_per_s": 256 / med * 1e3, "what": "K1,K2,masks,correspondences,dense maps,K3 per edit; inputs and stacks in HBM"}), flush=True)
            del levels, outs
        del eng


def config2_recorded_stack():
    """SURVEY.md 8(d) config 2, recorded-stack variant: the three recorded levels (1280,32^2), (640,64^2), (320,64^2) for
    all T = 50 timesteps of one edit warped through that edit's maps: 1.05 GB read + 1.05 GB written."""
    K = GuidedStableDiffuser.get_depth_intrinsics()
    depth, bg, mask = synthetic_scene(512, 0)
    eng = EditEngine(dev, 1, 512, 512)
    res = eng.run(*(torch.from_numpy(a).to(dev)[None].contiguous() for a in (depth, bg, mask)), K,
                  [make_rigid(30.0, [0.0, 1.0, 0.0], [0.3, 0.0, 0.2])], poisson=False)
    shapes = [(1280, 32), (640, 64), (320, 64)]
    T = 50
    maps1 = warp.dense_source_maps(res.corr, res.n_corr, 512, [s for _, s in shapes], res.winner_src)
    maps = [m.expand(T, -1).contiguous() for m in maps1]             # the same edit at every timestep
    g = torch.Generator(device=dev).manual_seed(5)
    stacks = [torch.randn((T, c, s, s), generator=g, device=dev) for c, s in shapes]
    outs = [torch.empty_like(a) for a in stacks]
    med, mn = timeit(lambda: warp.warp_stacks(stacks, maps, outs), n=20, warm=3)
    for a, m, o in zip(stacks, maps, outs):
        idx = m[7].long().clamp(min=0)
        assert torch.equal(o[7].flatten(1), a[7].flatten(1)[:, idx] * (m[7] >= 0))
    algo = 2 * sum(a.numel() for a in stacks) * 4 + sum(m.numel() for m in maps) * 4
    print(json.dumps({"config": "config2_recorded_stack", "timesteps": T, "ms_per_edit": med, "algorithmic_bytes": algo,
                      "achieved_gbs": algo / med / 1e6, "frac_of_hbm_peak": algo / med / 1e6 / PEAK}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "3", "4"]
    if "2" in which:
        config2_recorded_stack()
    if "1" in which:
        config1_and_5()
    if "3" in which:
        config3()
    if "4" in which:
        config4()
