#!/usr/bin/env python
"""bench.py - activation-stack warps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 2): per GPU, 256 independent edits, each with its own full SD2-depth
activation stack (64^2x320, 32^2x640, 16^2x1280, 8^2x1280 fp32 = 9.5 MB) warped through that edit's winner-index
map.  The 256 maps come from 256 real splats (16 synthetic depths x 16 rigid transforms, config 4 recipe).
One *step* = one pass of K3 over the 256 stacks (2.43 GB read + 2.43 GB written, far larger than L2).

  value     device-resident: stacks and maps already in HBM, one kernel launch per step, CUDA-event timed.
  e2e       the same 256 edits through the public batched API with HOST (pinned) buffers: H2D of depth, mask and
            stack, K1 -> K2 -> masks -> correspondences -> maps -> K3, D2H of the warped stack and counts.
  roofline  algorithmic bytes (19,027,200 B per warp) / K3 launch time vs the measured HBM copy peak.
  cpu_baseline  the oracle port of the reference path (NumPy geometry + torch CPU index gather) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

LEVELS = [(320, 64), (640, 32), (1280, 16), (1280, 8)]        # (channels, side) - BASELINE config 2
S = 512
EDITS_PER_GPU = 256
STACK_FLOATS = sum(c * s * s for c, s in LEVELS)                 # 2,375,680
ALGO_BYTES_PER_WARP = 2 * STACK_FLOATS * 4 + sum(s * s for _, s in LEVELS) * 4   # 19,027,200
METRIC = "activation_stack_warps_per_sec"
UNIT = "warps/s"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def edit_recipe(n_edits: int):
    """Config-4 recipe: 16 synthetic depths (seeds 100..115, disc radius U[60,160], centre jitter +-60 px) x 16
    transforms (angles linspace(-90,90,16) about (0,1,0); t = (0.1k-0.8, 0, 0.05k))."""
    rng = np.random.default_rng(1234)
    scenes = []
    for i in range(16):
        scenes.append(dict(S=S, seed=100 + i, cx=256.0 + float(rng.uniform(-60, 60)), cy=280.0 + float(rng.uniform(-60, 60)),
                           radius=float(rng.uniform(60, 160))))
    angles = np.linspace(-90.0, 90.0, 16)
    edits = []
    for e in range(n_edits):
        si, k = (e // 16) % 16, e % 16
        edits.append((si, float(angles[k]), (0.0, 1.0, 0.0), (0.1 * k - 0.8, 0.0, 0.05 * k)))
    return scenes, edits


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).

    The timed region is only tens of milliseconds, so the sampler is an in-process NVML thread (one query every
    ~1 ms, no start-up delay); `nvidia-smi -lms` is the fallback when the NVML binding is missing."""

    _REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.sm, self.mx, self.reasons, self._stop, self._thread, self.nvml = [], [], set(), False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if ids and all(v.strip().isdigit() for v in ids) and index < len(ids):
            return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        try:
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        except Exception:
            pass
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                bits = int(get_reasons(self.handle))
                for name, bit in self._REASONS:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.nvml is not None:
            self._thread = threading.Thread(target=self._sample_nvml, daemon=True)
            self._thread.start()
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self._thread is not None:
            self._stop = True
            self._thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 1 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference path)
# ------------------------------------------------------------------------------------------------
def cpu_warp_sample(n_stacks: int, with_geometry: bool, seed: int = 0, first_edit: int = 0):
    """Times the oracle port on `n_stacks` edits of the bench workload.  Returns (gather warps/s, e2e warps/s)."""
    from oracle import dh_oracle as O
    scenes, edits = edit_recipe(first_edit + n_stacks)
    edits = edits[first_edit:]
    K = O.get_depth_intrinsics()
    g = torch.Generator().manual_seed(seed)
    t_geo = t_gather = 0.0
    cache = {}
    for (si, angle, axis, t) in edits:
        if si not in cache:
            cache[si] = O.synthetic_scene(**scenes[si])
        depth, bg, mask = cache[si]
        t0 = time.perf_counter()
        o = O.transform_depth_pc(depth, bg, mask, K, angle, axis, tuple(float(np.float32(v)) for v in t), poisson=False)
        P = S * S
        ws = np.where(o["winner"] < 0, -1, np.where(o["winner"] < P, o["winner"], 0))
        fgw = o["winner"] >= P
        ws[fgw] = o["fg_index"][o["winner"][fgw] - P]
        maps = [O.dense_source_map(o["correspondences"], S, s, ws) for _, s in LEVELS]
        t_geo += time.perf_counter() - t0
        stack = [torch.randn((c, s, s), generator=g) for c, s in LEVELS]
        t0 = time.perf_counter()
        for A, m in zip(stack, maps):
            idx = torch.from_numpy(np.where(m >= 0, m, 0).astype(np.int64))
            out = A.flatten(1)[:, idx]                       # the reference's gather: feat_map[..., y, x] (losses.py:46-47)
            out[:, torch.from_numpy(m < 0)] = 0
        t_gather += time.perf_counter() - t0
    return n_stacks / t_gather, n_stacks / (t_gather + t_geo)


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))      # the CPU arm gets every core the container allows
    except OSError:
        pass
    n = 2
    try:
        pool = CpuPool()
        for _ in range(args.warmup):
            pool.rate(1)
        t0 = time.perf_counter()
        vals = [pool.rate(n) for _ in range(args.steps)]
        dt = time.perf_counter() - t0
        pool.close()
        cores = pool.workers
        sample = (f"per step: {n} edits on each of {cores} worker processes (oracle NumPy port of transform_depth_pc + dense maps + "
                  f"torch CPU index gather of the 4-level stack), geometry included; value = edits of the whole pool / wall time")
    except Exception as exc:                               # noqa: BLE001 - no process pool on this box: one process, all torch threads
        print(f"[bench] CPU pool unavailable ({exc}); timing one process", file=sys.stderr)
        for _ in range(args.warmup):
            cpu_warp_sample(1, True)
        t0 = time.perf_counter()
        vals = [cpu_warp_sample(2 * n, True)[1] for _ in range(args.steps)]
        dt = time.perf_counter() - t0
        cores = 1
        sample = (f"per step: {2 * n} edits in one process (oracle NumPy port of transform_depth_pc + dense maps + torch CPU index "
                  f"gather of the 4-level stack), geometry included")
    e2e = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": e2e, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: full SD2-depth activation stack (64^2x320,32^2x640,16^2x1280,8^2x1280) per edit, "
                                   "512^2 depth, config-4 edit recipe"},
            "cpu_baseline": {"value": e2e, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE configs 1, 3, 4, 5) reported under "variants" of the same JSON line
# ------------------------------------------------------------------------------------------------
def _median_ms(fn, dev, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize(dev)
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def variant_single_edits(dev, with_cpu: bool):
    """BASELINE config 1 (one 512^2 edit + one 64x64x320 activation) and config 5 (1024^2 stress A/B/C): device time of
    K1 -> K2 -> masks -> correspondences [-> source map -> K3], wall time of the public call with its count read-back, and the
    CPU time of the oracle port on the same edit (config 1 only)."""
    from diffusionhandles_b200 import warp
    from diffusionhandles_b200.synthetic import synthetic_scene
    from diffusionhandles_b200.engine import EditEngine, make_rigid
    from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
    K = GuidedStableDiffuser.get_depth_intrinsics()
    out = {}
    cases = (("config1", 512, dict(S=512, seed=0), 30.0, (0.3, 0.0, 0.2)),
             ("config5A", 1024, dict(S=1024, seed=0, cx=512.0, cy=560.0, radius=300.0), 90.0, (1.5, 0.0, 1.0)),
             ("config5B", 1024, dict(S=1024, seed=0, cx=512.0, cy=560.0, radius=300.0, quantize=0.1), 90.0, (1.5, 0.0, 1.0)),
             ("config5C", 1024, dict(S=1024, seed=0, cx=512.0, cy=560.0, radius=300.0), 60.0, (-2.0, 0.0, -1.5)))
    for name, S_, scene, angle, t in cases:
        depth, bg, mask = synthetic_scene(**scene)
        eng = EditEngine(dev, 1, S_, S_)
        td, tb, tm = (torch.from_numpy(a).to(dev)[None].contiguous() for a in (depth, bg, mask))
        rg = [make_rigid(angle, [0.0, 1.0, 0.0], list(t))]
        res = eng.run(td, tb, tm, K, rg, poisson=False)
        rec = {"S": S_, "n_fg": int(res.n_fg_host[0]), "n_corr": int(res.n_corr_host[0]),
               "gpu_ms": _median_ms(lambda: eng.run(td, tb, tm, K, rg, poisson=False, sync_counts=False), dev),
               "gpu_ms_with_hole_fill": _median_ms(lambda: eng.run(td, tb, tm, K, rg, poisson=True, sync_counts=False), dev)}
        # the same launch chain replayed from a CUDA graph (no host launch gaps)
        g, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            eng.run(td, tb, tm, K, rg, poisson=False, sync_counts=False)
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                eng.run(td, tb, tm, K, rg, poisson=False, sync_counts=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        rec["gpu_ms_graph"] = _median_ms(g.replay, dev)
        t0 = time.perf_counter()
        for _ in range(20):
            eng.run(td, tb, tm, K, rg, poisson=False, sync_counts=True)
        rec["wall_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        if name == "config1":
            A = torch.randn((1, 320, 64, 64), generator=torch.Generator(device=dev).manual_seed(1), device=dev)
            o = torch.empty_like(A)

            def edit_and_warp():
                r = eng.run(td, tb, tm, K, rg, poisson=False, sync_counts=False)
                warp.warp_stacks([A], warp.dense_source_maps(r.corr, r.n_corr, S_, [64], r.winner_src), [o])
            rec["gpu_ms_with_warp_320x64x64"] = _median_ms(edit_and_warp, dev)
            if with_cpu:
                from oracle import dh_oracle as O
                t0 = time.perf_counter()
                oc = O.transform_depth_pc(depth, bg, mask, O.get_depth_intrinsics(), angle, (0, 1, 0), tuple(float(np.float32(v)) for v in t),
                                          poisson=False)
                m = O.dense_source_map(oc["correspondences"], S_, 64)
                idx = torch.from_numpy(np.where(m >= 0, m, 0).astype(np.int64))
                _ = A[0].cpu().flatten(1)[:, idx]
                rec["cpu_ms"] = (time.perf_counter() - t0) * 1e3
                rec["cpu_kind"] = ("oracle port (vectorised NumPy; the reference's own Python-loop z-buffer takes ~0.7 s for this edit, "
                                   "BASELINE.md), 1 process")
                assert np.array_equal(res.correspondences(0).cpu().numpy(), oc["correspondences"]), "config-1 correspondences differ from the oracle"
        out[name] = rec
        del eng
    return out


def variant_set_foreground(dev):
    """DiffusionHandles.set_foreground at the real size (diffusion_handles.py:88-110; the reference's SuperLU solve takes 0.5-1.8 s
    on the bundled scenes): config-1 scene, 15 px dilated mask, fp64 CG on the device."""
    from diffusionhandles_b200.diffusion_handles import DiffusionHandles
    from diffusionhandles_b200 import depth_transform as dt
    from diffusionhandles_b200.synthetic import synthetic_scene
    depth, bg, mask = synthetic_scene(512, 0)
    td, tb, tm = (torch.from_numpy(a).to(dev)[None, None] for a in (depth, bg, mask))
    dh = DiffusionHandles()
    out = dh.set_foreground(td, tm, tb)
    iters = int(dt._poisson_device.last_iters[0])
    n_unknown = int((out[0, 0] != td[0, 0]).sum().item())
    t0 = time.perf_counter()
    for _ in range(5):
        dh.set_foreground(td, tm, tb)
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / 5 * 1e3
    return {"set_foreground_512": {"unknowns_approx": n_unknown, "cg_iterations": iters, "converged": iters > 0, "wall_ms": wall,
                                   "reference": "scipy SuperLU on the CPU: 0.5-1.8 s per bundled scene (tests/golden/photogen_ref.json)"}}


def variant_guidance_loss(dev):
    """BASELINE config 3: guidance loss forward + backward on the recorded-stack shapes for the 50 recorded timesteps.
    kernels: 50 evaluations (one per timestep, 3.1 GB of distinct inputs -> L2-cold) captured in one CUDA graph;
    api: the same evaluations through guidance_loss + torch.autograd.grad."""
    from diffusionhandles_b200 import losses
    from diffusionhandles_b200.synthetic import synthetic_scene
    from diffusionhandles_b200.engine import EditEngine, make_rigid
    from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser, make_guidance_weight_schedule
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
    curs = [torch.randn((T, c, s, s), generator=g3, device=dev) for c, s in shapes]
    algo = 3 * sum(c * s * s for c, s in shapes) * 4
    plan = losses._plan_for(pc, 64, dev)

    def kernels(t):
        losses._launch([c[t] for c in curs], [o[t] for o in origs], [True] * 3, [1.0] * 3, [1.0] * 3, plan, 1, 1)
    kernels(0)
    torch.cuda.synchronize(dev)
    gr, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        kernels(1)
        side.synchronize()
        with torch.cuda.graph(gr, stream=side):
            for t in range(T):
                kernels(t)
    torch.cuda.current_stream(dev).wait_stream(side)
    k_ms = _median_ms(gr.replay, dev, n=10, warm=2) / T
    state = {"i": 0}

    def api():
        t = state["i"] % T
        state["i"] += 1
        cs = [c[t].requires_grad_(True) for c in curs]
        total, _ = losses.guidance_loss(cs, [o[t] for o in origs], pc, [1.0] * 3, [1.0] * 3)
        torch.autograd.grad(total, cs)
    for _ in range(6):
        api()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(150):
        api()
    torch.cuda.synchronize(dev)
    api_ms = (time.perf_counter() - t0) / 150 * 1e3

    def api_direct():       # value + gradients without an autograd node (what guided_denoise calls)
        t = state["i"] % T
        state["i"] += 1
        losses.guidance_loss_and_grad([c[t] for c in curs], [o[t] for o in origs], pc, [1.0] * 3, [1.0] * 3)
    for _ in range(6):
        api_direct()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(150):
        api_direct()
    torch.cuda.synchronize(dev)
    api_direct_ms = (time.perf_counter() - t0) / 150 * 1e3
    # the weights guided_inference really uses (layer 0 always has weight 0 and is skipped; steps alternate between the layers)
    sched = make_guidance_weight_schedule(1.5, 1.25)
    gr2 = torch.cuda.CUDAGraph()
    n_eval = 0
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr2, stream=side):
            for t in range(38):
                fgw, bgw = sched(t, 0)
                keep = [i for i in range(3) if fgw[i] or bgw[i]]
                if keep:
                    losses._launch([curs[i][t] for i in keep], [origs[i][t] for i in keep], [True] * len(keep), [fgw[i] for i in keep],
                                   [bgw[i] for i in keep], plan, 1, 1)
                    n_eval += 1
    torch.cuda.current_stream(dev).wait_stream(side)
    sched_ms = _median_ms(gr2.replay, dev, n=10, warm=2) / max(n_eval, 1)
    peak, _ = measured_peak_hbm()
    return {"config3": {"evaluations_timed": 150, "n_corr": int(res.n_corr_host[0]), "ms_per_evaluation_kernels": k_ms,
                        "ms_per_evaluation_api": api_ms, "ms_per_evaluation_api_direct": api_direct_ms,
                        "api": "api = losses.guidance_loss + torch.autograd.grad (a Python autograd node); api_direct = "
                               "losses.guidance_loss_and_grad (value and gradients from the same launch, no autograd node - the call "
                               "guided_inference makes); wall clock of 150 back-to-back evaluations",
                        "algorithmic_bytes": algo, "achieved_gbs": algo / k_ms / 1e6,
                        "frac": algo / k_ms / 1e6 / peak, "l2": "cold: 50 timesteps of distinct inputs (3.1 GB) per replay",
                        "layers": "all three recorded layers active (1280x32^2, 640x64^2, 320x64^2), global_avg background, patch 1",
                        "ms_per_evaluation_kernels_reference_schedule": sched_ms,
                        "reference_schedule": "weights of guided_stable_diffuser.py:336-373: zero-weight layers (layer 0 always) skipped"}}


def variant_guided_step(dev):
    """the elementwise steps around the U-Net on SD2-depth latents (1,4,64,64) - the eager torch ops of
    guided_stable_diffuser.py:434 and :470-474 (DDIMScheduler.step) against the one-launch kernels.  Wall clock of 200
    back-to-back steps (launch bound: that is the cost the fusion removes) and device time of the same steps replayed from a
    CUDA graph."""
    from diffusionhandles_b200.guided_loop import DDIMSchedule, cfg_ddim_step, latent_step
    g = torch.Generator(device=dev).manual_seed(5)
    lat, grad, nu, nt = (torch.randn((1, 4, 64, 64), generator=g, device=dev) for _ in range(4))
    sched = DDIMSchedule()
    sched.set_timesteps(50)
    t = 500
    a_t, a_p = sched.alphas_cumprod[t], sched.alphas_cumprod[t - 20]
    co = sched.coefficients(t)

    def eager():
        x = lat - grad * 0.1
        eps = nu + 7.5 * (nt - nu)
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_p ** 0.5 * x0 + (1 - a_p - 0.0 ** 2) ** 0.5 * eps

    def fused():
        return cfg_ddim_step(nu, nt, latent_step(lat, grad, 0.1), co)
    same = bool(torch.equal(eager(), fused()))
    out = {}
    for name, fn in (("eager", eager), ("fused", fused)):
        for _ in range(20):
            fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(200):
            fn()
        torch.cuda.synchronize(dev)
        out[f"{name}_wall_us"] = (time.perf_counter() - t0) / 200 * 1e6
        gr, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fn()
            side.synchronize()
            with torch.cuda.graph(gr, stream=side):
                for _ in range(50):
                    fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        out[f"{name}_device_us"] = _median_ms(gr.replay, dev, n=10, warm=2) / 50 * 1e3
    return {"guided_step": {**out, "bit_identical_to_eager": same, "launches": {"eager": 11, "fused": 2},
                            "what": "latents - 0.1 grad, CFG 7.5 combine and the DDIM update (eta 0) on (1,4,64,64) fp32 latents, per denoising step"}}


def variant_strong_scaling(dev, rank, world, steps=20):
    """BASELINE config 4 as written: 256 (depth, transform) edits IN TOTAL, edit e on rank e mod N, the whole device-resident
    pipeline per edit (K1 -> K2 -> masks -> correspondences -> dense maps -> K3 on the edit's own config-2 stack) and the ONE
    NCCL gather of the per-edit result records inside the timed region.  Device time, max over ranks."""
    import torch.distributed as dist
    from diffusionhandles_b200.batch import DeviceSweep, gather_records, shard_edits
    from diffusionhandles_b200.synthetic import synthetic_scene
    from diffusionhandles_b200.engine import make_rigid
    from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
    total = 256
    scenes, edits = edit_recipe(total)
    mine = shard_edits(total, rank, world)
    K = GuidedStableDiffuser.get_depth_intrinsics()
    sc = {}
    for e in mine:
        si = edits[e][0]
        if si not in sc:
            sc[si] = [torch.from_numpy(a).to(dev) for a in synthetic_scene(**scenes[si])]
    depth = torch.stack([sc[edits[e][0]][0] for e in mine]).contiguous()
    bg = torch.stack([sc[edits[e][0]][1] for e in mine]).contiguous()
    mask = torch.stack([sc[edits[e][0]][2] for e in mine]).contiguous()
    rigids = [make_rigid(edits[e][1], list(edits[e][2]), list(edits[e][3])) for e in mine]
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    levels = [torch.randn((len(mine), c, s, s), generator=gen, dtype=torch.float32, device=dev) for c, s in LEVELS]
    n_local = len(mine)
    chunk = n_local if n_local <= 64 else 64
    # (parallel graph branches for small shards were a gain before the geometry kernels were re-tiled and are a loss now:
    # tools/sweep_branches.py; one chain per chunk)
    branches = int(os.environ.get("DH_BENCH_BRANCHES", "1"))
    sweep = DeviceSweep(dev, S, LEVELS, depth, bg, mask, K, rigids, levels, chunk=chunk, use_graph=True, branches=branches)
    sweep.capture()
    # the replayed graphs must give what the plain launch chain gives (first chunk: counts and the warped stack)
    sweep.run()
    ref = DeviceSweep(dev, S, LEVELS, depth[:chunk], bg[:chunk], mask[:chunk], K, rigids[:chunk], [l[:chunk] for l in levels],
                      chunk=chunk, use_graph=False)
    ref.run()
    torch.cuda.synchronize(dev)
    assert torch.equal(sweep.n_corr[:chunk], ref.n_corr) and int(ref.n_corr.sum()) > 0, "graph replay differs from the plain launch chain"
    assert all(torch.equal(a[:chunk], b) for a, b in zip(sweep.outs, ref.outs)), "graph replay: warped stacks differ"
    del ref

    def step():
        sweep.run()
        rec = sweep.records()
        return gather_records(rec, dst=0) if world > 1 else rec
    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        out = step()
    ev1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([ev0.elapsed_time(ev1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n_corr_sum = int(out[:, 0].sum().item()) if rank == 0 else None
    del sweep
    return {"config4_strong": {"edits_total": total, "edits_per_rank": n_local, "chunk": chunk, "graph_branches": branches, "ms_per_sweep": ms,
                               "edits_per_s": total / ms * 1e3, "n_corr_sum": n_corr_sum,
                               "what": "K1,K2,masks,correspondences,dense maps,K3 per edit (device-resident, CUDA-graph replay per chunk) + "
                                       "one NCCL gather of the result records; device time, max over ranks",
                               "scaling": "strong"}}

# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_workload(dev, n_edits: int, full_map: bool):
    """Runs the real geometry (K1 -> K2 -> masks -> correspondences) for n_edits edits and returns the per-level
    source maps (device) plus host copies of the geometry inputs."""
    from diffusionhandles_b200 import warp
    from diffusionhandles_b200.synthetic import synthetic_scene
    from diffusionhandles_b200.engine import EditEngine, make_rigid
    from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
    scenes, edits = edit_recipe(n_edits)
    K = GuidedStableDiffuser.get_depth_intrinsics()
    chunk = 16
    eng = EditEngine(dev, chunk, S, S)
    sc = [synthetic_scene(**s) for s in scenes]
    depth_h = torch.empty((n_edits, S, S), dtype=torch.float32).pin_memory()
    bg_h = torch.empty((n_edits, S, S), dtype=torch.float32).pin_memory()
    mask_h = torch.empty((n_edits, S, S), dtype=torch.float32).pin_memory()
    rigids = []
    for e, (si, angle, axis, t) in enumerate(edits):
        depth_h[e] = torch.from_numpy(sc[si][0]); bg_h[e] = torch.from_numpy(sc[si][1]); mask_h[e] = torch.from_numpy(sc[si][2])
        rigids.append(make_rigid(angle, list(axis), list(t)))
    # the same inputs as a scene table (16 scenes) + a scene index per edit: what the e2e leg uploads
    scene_h = [torch.empty((len(sc), S, S), dtype=torch.float32).pin_memory() for _ in range(3)]
    for i, arrs in enumerate(sc):
        for k in range(3):
            scene_h[k][i] = torch.from_numpy(arrs[k])
    scene_index = [e[0] for e in edits]
    maps = [torch.empty((n_edits, s * s), dtype=torch.int32, device=dev) for _, s in LEVELS]
    n_corr = torch.empty(n_edits, dtype=torch.int32, device=dev)
    for e0 in range(0, n_edits, chunk):
        e1 = e0 + chunk
        res = eng.run(depth_h[e0:e1].to(dev), bg_h[e0:e1].to(dev), mask_h[e0:e1].to(dev), K, rigids[e0:e1], poisson=False)
        ms = warp.dense_source_maps(res.corr, res.n_corr, S, [s for _, s in LEVELS], res.winner_src if full_map else None)
        for dst, m in zip(maps, ms):
            dst[e0:e1] = m
        n_corr[e0:e1] = res.n_corr
    torch.cuda.synchronize(dev)
    del eng
    return dict(depth_h=depth_h, bg_h=bg_h, mask_h=mask_h, rigids=rigids, K=K, maps=maps, n_corr=n_corr, scene_h=scene_h,
                scene_index=scene_index)


def time_kernel_steps(fn, steps: int, warmup: int, dev):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    return [a.elapsed_time(b) for a, b in evs]


def run_ours(args):
    import torch.distributed as dist
    from diffusionhandles_b200 import warp, _native
    from diffusionhandles_b200.batch import EditWarpPipeline, bind_host_to_gpu_numa_node
    rank, world, local = dist_env()
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _native.load()
    if os.environ.get("DH_BENCH_ONLY_STRONG"):          # developer shortcut: only the strong-scaling variant
        v = variant_strong_scaling(dev, rank, world)
        if rank == 0:
            print(json.dumps(v), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # host threads + pinned staging buffers on the GPU's own NUMA node (matters for the e2e leg at N > 1)
    full_affinity = os.sched_getaffinity(0)
    host_binding = bind_host_to_gpu_numa_node(dev) if os.environ.get("DH_BENCH_NUMA_BIND", "1") == "1" else None
    n_edits = EDITS_PER_GPU
    wl = build_workload(dev, n_edits, full_map=True)
    gen = torch.Generator(device=dev).manual_seed(2 + rank)
    levels = [torch.randn((n_edits, c, s, s), generator=gen, dtype=torch.float32, device=dev) for c, s in LEVELS]
    outs = [torch.empty_like(l) for l in levels]
    maps = wl["maps"]

    def step():
        warp.warp_stacks(levels, maps, outs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident metric: K steps, barrier + sync on both sides, max over ranks ----
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n_edits * args.steps / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream ----
    per_launch = time_kernel_steps(step, max(args.steps, 100), 3, dev)   # >= 100 launches: ~75 ms under load for the clock sampler
    clocks = sampler.stop()          # sampled over the K timed steps and the per-launch roofline loop
    k3_ms = float(np.mean(per_launch))
    achieved = ALGO_BYTES_PER_WARP * n_edits / (k3_ms * 1e-3) / 1e9
    peak, peak_kind = measured_peak_hbm()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k3_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile": True, "ms_per_launch": k3_ms, "achieved_gbs": achieved, "frac": achieved / peak}), flush=True)
        return

    # parity spot check inside the bench: the timed kernel really produced the gather (3 stacks, torch indexing)
    for l, m, o in zip(levels, maps, outs):
        for e in (0, n_edits // 2, n_edits - 1):
            idx = m[e].long().clamp(min=0)
            ref = l[e].flatten(1)[:, idx] * (m[e] >= 0)
            assert torch.equal(o[e].flatten(1), ref), "K3 output differs from an index gather"

    # ---- variant: correspondence-only maps (SURVEY.md row 9 dense form; ~80% of the cells are empty) ----
    wl2 = build_workload(dev, n_edits, full_map=False)
    maps2 = wl2["maps"]
    outs2 = [torch.empty_like(l) for l in levels]
    corr_only = time_kernel_steps(lambda: warp.warp_stacks(levels, maps2, outs2), 10, 3, dev)
    coverage = float(np.mean([(m >= 0).float().mean().item() for m in maps]))
    coverage2 = float(np.mean([(m >= 0).float().mean().item() for m in maps2]))
    del wl2, maps2, outs2

    # ---- end to end through the public batched API with pinned HOST buffers ----
    e2e_steps = max(2, min(args.steps, 5))
    pipe = EditWarpPipeline(dev, S, LEVELS, chunk=16, n_streams=4, full_winner_map=True)      # (tools/tune_e2e.py sweep)
    levels_h = [torch.empty((n_edits, c, s, s), dtype=torch.float32).pin_memory() for c, s in LEVELS]
    for h, d in zip(levels_h, levels):
        h.copy_(d)
    outs_h = [torch.empty((n_edits, c, s, s), dtype=torch.float32).pin_memory() for c, s in LEVELS]
    n_corr_h = torch.empty(n_edits, dtype=torch.int32).pin_memory()

    def e2e_step():       # every scene's depth / background / mask crosses PCIe once (scene table), every stack once each way
        pipe.run_host(wl["scene_h"][0], wl["scene_h"][1], wl["scene_h"][2], wl["K"], wl["rigids"], levels_h, outs_h, n_corr_h,
                      scene_index=wl["scene_index"])
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_edits * e2e_steps / float(te.item())
    assert torch.equal(n_corr_h, wl["n_corr"].cpu()), "e2e correspondences differ from the device-resident run"
    for h, o in zip(outs_h, outs):
        assert torch.equal(h[-1], o[-1].cpu()), "e2e warped stack differs from the device-resident run"

    # ---- one NCCL gather of small per-edit result records (not on the hot path) ----
    gather_ms = None
    if world > 1:
        from diffusionhandles_b200.batch import gather_records
        rec = torch.stack([wl["n_corr"].to(torch.int64), outs[0].flatten(1).sum(1).double().view(torch.int64)], dim=1).contiguous()
        gather_records(rec, dst=0)      # first call sets up the NCCL communicator
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        gathered = gather_records(rec, dst=0)
        torch.cuda.synchronize(dev)
        assert rank != 0 or gathered.shape[0] == world * n_edits
        gather_ms = 1e3 * (time.perf_counter() - t0)

    # ---- secondary workloads: config 4 as written (strong scaling, every N), configs 1 / 3 / 5 (N = 1 only) ----
    h2d_per_step = int(pipe.h2d_bytes_per_edit(edits_per_scene=16) * n_edits)
    d2h_per_step = pipe.d2h_bytes_per_edit() * n_edits
    del pipe, levels_h, outs_h
    variants = {}
    if not os.environ.get("DH_BENCH_SKIP_VARIANTS"):
        del levels, outs
        torch.cuda.empty_cache()
        variants.update(variant_strong_scaling(dev, rank, world))
        if world == 1:
            variants.update(variant_guidance_loss(dev))
            variants.update(variant_single_edits(dev, with_cpu=True))
            variants.update(variant_set_foreground(dev))
            try:
                variants.update(variant_guided_step(dev))
            except Exception as exc:                   # noqa: BLE001 - a side measurement must not cost the headline line
                variants["guided_step"] = {"error": repr(exc)}
        torch.cuda.empty_cache()

    line = None
    if rank == 0:
        os.sched_setaffinity(0, full_affinity)        # the CPU baseline gets every host core again
        cpu_warp_sample(1, True)
        cpu_gather, cpu_serial = cpu_warp_sample(8, True)
        cpu_cores, cpu_e2e, per_worker = 1, cpu_serial, 0
        if world == 1:                                # BASELINE: the CPU leg runs on rank 0 at N=1 only
            try:
                pool = CpuPool()
                per_worker = 4
                cpu_e2e, cpu_cores = pool.rate(per_worker), pool.workers
                pool.close()
            except Exception as exc:                   # noqa: BLE001 - the serial figure stands
                print(f"[bench] CPU pool unavailable ({exc}); reporting the single-process figure", file=sys.stderr)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: full SD2-depth activation stack (64^2x320,32^2x640,16^2x1280,8^2x1280) per edit, "
                                   "256 edits per GPU, each through its own winner-index map (512^2 splat, config-4 edit recipe)",
                       "edits_per_gpu": n_edits, "bytes_per_warp": ALGO_BYTES_PER_WARP, "map": "full winner-index map",
                       "map_coverage": coverage, "l2": "inputs (2.43 GB read + 2.43 GB written per step) far larger than L2; no flush",
                       "parallelism": f"{world} x independent shards, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles/k3_traffic.json (one ncu --set full capture of this kernel and "
                                                              "workload, dram__bytes_read.sum + dram__bytes_write.sum per launch; not re-measured in this run)",
                         "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                         "kernel": "warp_dense_tma_kernel", "ms_per_launch": k3_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_WARP * n_edits},
            "cpu_baseline": {"value": cpu_e2e, "unit": UNIT, "cores": cpu_cores, "kind": "port",
                             "sample": (f"{per_worker} edits on each of {cpu_cores} worker processes" if per_worker else
                                        "8 edits in one process (the all-core CPU leg runs at N=1 and in --impl reference)") +
                                       " of the same workload: oracle NumPy port of transform_depth_pc + dense maps + torch CPU index "
                                       f"gather of the 4-level stack (one process alone: {cpu_serial:.1f} warps/s; its gather alone: "
                                       f"{cpu_gather:.1f} warps/s)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h_per_step, "steps": e2e_steps,
                    "host_numa_binding": host_binding,
                    "what": "pinned host scene table (16 scenes: depth, background, mask - uploaded once per scene) + one stack per edit "
                            "-> K1,K2,masks,correspondences,maps,K3 -> pinned host warped stack"},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "variants": {"corr_only_map": {"ms_per_launch": float(np.mean(corr_only)), "map_coverage": coverage2,
                                           "warps_per_s": n_edits / (float(np.mean(corr_only)) * 1e-3)}, **variants},
            "wall_s_timed_region": wall, "result_gather_ms": gather_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _cpu_worker(job):
    """One host process of the CPU pool: `n` edits starting at `first`, single-threaded torch."""
    first, n = job
    torch.set_num_threads(1)
    cpu_warp_sample(n, True, seed=first, first_edit=first)
    return n


class CpuPool:
    """The reference path is single-threaded Python/NumPy, but edits are independent: to give the CPU arm every
    host thread, one worker process per available core each runs its own edits (spawned, so no CUDA/OpenMP state
    is inherited).  `rate(k)` = edits/s of the whole pool on k edits per worker."""

    def __init__(self, workers: int = 0):
        import multiprocessing as mp
        self.workers = workers or min(len(os.sched_getaffinity(0)), 64)     # bounded: each worker imports torch
        self.pool = mp.get_context("spawn").Pool(self.workers)
        self.pool.map(_cpu_worker, [(0, 1)] * self.workers)           # import + first-call costs outside the timing

    def rate(self, edits_per_worker: int) -> float:
        jobs = [(w * edits_per_worker, edits_per_worker) for w in range(self.workers)]
        t0 = time.perf_counter()
        done = sum(self.pool.map(_cpu_worker, jobs, chunksize=1))
        return done / (time.perf_counter() - t0)

    def close(self):
        self.pool.terminate()
        self.pool.join()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--profile", action="store_true", help="device-resident K3 steps only (for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
