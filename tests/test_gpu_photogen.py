"""GPU (-m gpu): the CUDA path on the reference's OWN fixtures - the 20 bundled photogen scenes (ZoeDepth EXR maps, real
masks, tests/golden/photogen_inputs.npz) and all 90 edits of their transforms.json - through the public API
(DiffusionHandles.set_foreground, depth_transform.transform_depth) and the C ABI.  Correspondences are compared bit for
bit (values and order) with the SHA-256 the REAL reference produced (tests/golden/photogen_ref.json); every intermediate is
compared with the oracle, which the CPU suite pins to the same reference outputs."""
import numpy as np
import pytest
import torch

from oracle import dh_oracle as O
from helpers import sha, f32_translation, load_photogen, photogen_filled_bg
from test_gpu_parity import run_edit, compare_edit, K_NP

pytestmark = pytest.mark.gpu

META, GET = load_photogen()
SCENES = sorted(META)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def K(dev):
    from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
    return GuidedStableDiffuser.get_depth_intrinsics(device=dev)


@pytest.mark.parametrize("scene", SCENES)
def test_set_foreground_real_scene(dev, scene):
    """diffusion_handles.py:88-110 at the real size: 512^2, 15 px dilated mask, 12k-95k unknowns; the reference solves with
    SuperLU (0.5-1.8 s), here fp64 CG on the device.  Tolerance: 1e-4 relative to the depth range of the scene."""
    from diffusionhandles_b200.diffusion_handles import DiffusionHandles
    depth, bg, mask = GET(scene)
    ref = photogen_filled_bg(scene, GET)
    td, tb, tm = (torch.from_numpy(a).to(dev)[None, None] for a in (depth, bg, mask))
    out = DiffusionHandles().set_foreground(td, tm, tb)
    assert out.shape == (1, 1, 512, 512) and out.dtype == torch.float32
    got = out[0, 0].cpu().numpy()
    import scipy.ndimage
    dil = scipy.ndimage.binary_dilation(mask, iterations=15)
    assert np.array_equal(got[~dil], depth[~dil])                                   # outside the hole: the input depth, untouched
    tol = 1e-4 * float(depth.max() - depth.min())
    assert np.abs(got - ref).max() <= tol
    rows, vals = GET.rows(scene)                                                    # and against the stored reference rows
    assert np.abs(got[rows] - vals).max() <= tol


@pytest.mark.parametrize("scene", SCENES)
def test_transform_depth_real_scene(dev, K, scene):
    """Every edit of the scene's transforms.json: engine intermediates == oracle, correspondences == the reference's SHA."""
    from diffusionhandles_b200 import depth_transform as dt
    depth, bg, mask = GET(scene)
    bg2 = photogen_filled_bg(scene, GET)
    td, tb, tm = (torch.from_numpy(a).to(dev)[None, None] for a in (depth, bg2, mask))
    for name, e in META[scene]["edits"].items():
        t = f32_translation(e["translation"])
        o = O.transform_depth_pc(depth, bg2, mask, K_NP, e["rotation_angle"], e["rotation_axis"], t)
        assert sha(o["correspondences"]) == e["sha_corr"]                            # the oracle is the reference here
        eng, res = run_edit(dev, K, depth, bg2, mask, e["rotation_angle"], e["rotation_axis"], e["translation"])
        compare_edit(eng, res, o, 512)
        # the public call, as DiffusionHandles.transform_foreground makes it (diffusion_handles.py:143-159)
        disp, corr = dt.transform_depth(td, tb, tm, K, rot_angle=e["rotation_angle"], rot_axis=torch.tensor(e["rotation_axis"]),
                                        translation=torch.tensor(e["translation"]))
        assert corr.shape[0] == e["n_corr"] and sha(corr.numpy()) == e["sha_corr"], name
        assert np.abs(disp[0, 0].cpu().numpy() - o["disparity"]).max() <= 1e-3        # 0..255 scale, CG vs SuperLU
