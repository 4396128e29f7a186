"""Randomised parity sweep: CUDA path vs the oracle on random scenes / transforms (bit-exact comparison of the splat winners,
depth maps, masks, correspondences).  python tests/fuzz/fuzz_parity.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dh_oracle as O                                        # noqa: E402
from diffusionhandles_b200.engine import get_engine, make_rigid          # noqa: E402
from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser   # noqa: E402


def run(n_cases: int, seed: int, verbose: bool = True) -> int:
    """Returns the number of mismatching cases."""
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    K = GuidedStableDiffuser.get_depth_intrinsics()
    K_NP = K.numpy()
    bad = 0
    for case in range(n_cases):
        S = int(rng.choice([64, 96, 128, 160, 256, 320, 512]))
        kind = "smooth" if rng.random() < 0.4 else "disc"
        scene_seed = int(rng.integers(0, 10_000))
        if kind == "disc":
            scene = dict(S=S, seed=scene_seed, cx=float(rng.uniform(0.2, 0.8) * S), cy=float(rng.uniform(0.2, 0.8) * S),
                         radius=float(rng.uniform(0.03, 0.45) * S), quantize=float(rng.choice([0.0, 0.0, 0.1, 0.05])) or None)
        else:
            scene = dict(S=S, seed=scene_seed, kind="smooth")
        depth, bg, mask = O.synthetic_scene(**scene)
        axis_kind = rng.integers(0, 4)
        axis = [(0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), tuple(rng.normal(size=3).tolist())][axis_kind]
        angle = float(rng.uniform(-180, 180))
        big = rng.random() < 0.25
        t = tuple(float(v) for v in (rng.normal(size=3) * (2.5 if big else 0.3)))      # big: off-screen / behind the camera
        t32 = tuple(float(np.float32(v)) for v in t)
        norm = bool(rng.random() < 0.3)
        o = O.transform_depth_pc(depth, bg, mask, K_NP, angle, axis, t32, use_input_depth_normalization=norm, poisson=False)
        eng = get_engine(dev, 1, S, S, keep_points=True)
        td, tb, tm = (torch.from_numpy(a).to(dev)[None].contiguous() for a in (depth, bg, mask))
        res = eng.run(td, tb, tm, K, [make_rigid(angle, torch.tensor(axis, dtype=torch.float32), torch.tensor(t, dtype=torch.float32))],
                      use_input_depth_normalization=norm, poisson=False)
        P = S * S
        n_fg = int(res.n_fg_host[0])
        checks = {
            "n_fg": n_fg == len(o["fg_index"]),
            "centroid": np.array_equal(res.centroid[0].cpu().numpy(), o["centroid"]) if n_fg else True,
            "points": np.array_equal(res.points[0, : P + n_fg].cpu().numpy(), o["points"]),
            "pix": np.array_equal(res.pix[0, : P + n_fg].cpu().numpy().astype(np.int64), o["pix"]),
            "winner": np.array_equal(res.winner[0].cpu().numpy().astype(np.int64), o["winner"]),
            "depth_map": np.array_equal(res.depth_map[0].cpu().numpy(), o["depth_map"]),
            "target_mask": np.array_equal(res.target_mask[0].cpu().numpy().astype(bool), o["target_mask"]),
            "cleaned": np.array_equal(eng.unpack_bits(res.cleaned_bits)[0].cpu().numpy().astype(bool), o["cleaned"]),
            "corr": np.array_equal(res.correspondences(0).cpu().numpy(), o["correspondences"]),
            "disparity_raw": np.array_equal(res.disparity_raw[0].cpu().numpy(), o["disparity_raw"], equal_nan=True),
        }
        failed = [k for k, v in checks.items() if not v]
        if failed:
            bad += 1
            if verbose:
                print(f"MISMATCH case {case}: scene={scene} axis={axis} angle={angle} t={t} norm={norm}: {failed}", flush=True)
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    t0 = time.time()
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(f"{n} cases, {bad} mismatching, {time.time() - t0:.0f} s")
