"""CPU: the C-ABI library loads and exports every symbol include/dh_b200.h declares (no compute calls),
plus the host-side logic of the Python mirror."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import dh_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "dh_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from diffusionhandles_b200 import _native as N
    lib = N.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dh_b200.h but not exported"
    assert sorted(N.DECLARED_SYMBOLS) == syms, "ctypes signature table and header disagree"
    assert lib.dh_abi_version() == 1
    assert lib.dh_status_string(0) == b"ok" and lib.dh_status_string(-4) == b"workspace too small"


def test_library_is_sm100a_only():
    import subprocess
    from diffusionhandles_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_linspace_matches_torch():
    from diffusionhandles_b200 import _native as N
    lib = N.load()
    for (a, b, n) in ((-1.0, 1.0, 512), (-1.0, 1.0, 1024), (-1.0, 1.0, 333), (-47 / 79, 47 / 79, 48), (-1.0, 1.0, 2)):
        buf = (ctypes.c_float * n)()
        assert lib.dh_linspace_f32_host(a, b, n, buf) == 0
        assert np.array_equal(np.frombuffer(buf, dtype=np.float32), torch.linspace(a, b, n).numpy()), (a, b, n)
    assert lib.dh_linspace_f32_host(0.0, 1.0, 0, (ctypes.c_float * 1)()) == -1


def test_invalid_arguments_do_not_touch_the_gpu():
    from diffusionhandles_b200 import _native as N
    lib = N.load()
    assert lib.dh_unproject(None, 1, 8, 8, None, None, None, None, None) == -1
    assert lib.dh_warp_gather_dense(None, 1, 1, None) == -1
    assert lib.dh_splat_zbuffer(None, None, None, 0, 0, 0, 1, 1, None, None, None) == -1
    assert lib.dh_edit_workspace_bytes(1, 512, 512) > 3 * 4 * 512 * 512
    assert lib.dh_edit_workspace_bytes(0, 512, 512) == 0
    with pytest.raises(ValueError):
        N.check(-1, "x")
    with pytest.raises(RuntimeError):
        N.check(-2, "x")


def test_ellipse_rows_match_opencv_and_oracle():
    from diffusionhandles_b200.engine import ellipse_rows
    cv2 = pytest.importorskip("cv2")
    for k in range(1, 33):
        el = np.array([[(r >> j) & 1 for j in range(k)] for r in ellipse_rows(k)], np.uint8)
        assert np.array_equal(el, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))), k
        assert np.array_equal(el, O.ellipse_element(k))
    with pytest.raises(ValueError):
        ellipse_rows(33)


def test_make_rigid_follows_reference_host_arithmetic():
    from diffusionhandles_b200.engine import make_rigid
    rg = make_rigid(30.0, torch.tensor([0.0, 2.5, 0.0]), torch.tensor([0.3, 0.0, 0.2]))
    assert list(rg.axis) == [0.0, 1.0, 0.0]
    assert rg.cos_t == float(np.cos(np.radians(30.0))) and rg.sin_t == float(np.sin(np.radians(30.0)))
    assert list(rg.t) == [float(np.float32(0.3)), 0.0, float(np.float32(0.2))]
    a = O.normalize_axis([0.3, 0.9, -0.2])
    rg = make_rigid(torch.tensor(12.5), [0.3, 0.9, -0.2], [1, 2, 3])
    assert np.array_equal(np.array(list(rg.axis), np.float32), a)


def test_schedule_and_utils_match_oracle():
    from diffusionhandles_b200.guided_stable_diffuser import make_guidance_weight_schedule, GuidedStableDiffuser
    from diffusionhandles_b200.utils import pack_correspondences, unpack_correspondences
    for kind in ("constant", "linear", "quadratic"):
        s, o = make_guidance_weight_schedule(1.5, 1.25, 38, kind), O.guidance_weight_schedule(1.5, 1.25, 38, kind)
        for t in (0, 1, 2, 17, 37, 38, 49):
            for it in (0, 1, 2, 3, 5):
                assert s(t, it) == o(t, it)
    with pytest.raises(ValueError):
        make_guidance_weight_schedule(1, 1, 38, "cubic")
    assert np.array_equal(GuidedStableDiffuser.get_depth_intrinsics().numpy(), O.get_depth_intrinsics())
    c = pack_correspondences(*[torch.arange(5) + i for i in range(4)])
    assert c.shape == (5, 4) and [t.shape for t in unpack_correspondences(c)] == [(5, 1)] * 4


def test_cpu_tensors_fail_loudly():
    from diffusionhandles_b200 import depth_transform as dt, losses, _native as N
    with pytest.raises(N.NativeLibraryError):
        dt.depth_to_world_coords(torch.ones(1, 1, 8, 8), torch.eye(3))
    with pytest.raises(N.NativeLibraryError):
        pc = O.process_correspondences(np.zeros((0, 4), np.int64), 512)
        losses.compute_foreground_loss(torch.ones(2, 64, 64), torch.ones(2, 64, 64), pc, 1, (64, 64))


def test_product_and_oracle_scene_generators_agree():
    from diffusionhandles_b200.synthetic import synthetic_scene
    for kw in (dict(S=64, seed=3), dict(S=128, seed=5, cx=40.0, cy=70.0, radius=30.0, quantize=0.1)):
        for a, b in zip(synthetic_scene(**kw), O.synthetic_scene(**kw)):
            assert np.array_equal(a, b)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under diffusionhandles_b200/ may import or mention it, and the only other
    files that do are tests/, __graft_entry__.py (smoke) and bench.py (CPU legs)."""
    pkg = os.path.join(ROOT, "diffusionhandles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
    allowed = {"bench.py", "__graft_entry__.py"}
    for f in os.listdir(ROOT):
        if f.endswith(".py") and f not in allowed:
            with open(os.path.join(ROOT, f)) as fh:
                assert not re.search(r"^\s*(from|import)\s+oracle\b", fh.read(), flags=re.M), f"{f} imports the oracle"
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            with open(os.path.join(ROOT, "tools", f)) as fh:
                assert not re.search(r"^\s*(from|import)\s+oracle\b", fh.read(), flags=re.M), f"tools/{f} imports the oracle"
