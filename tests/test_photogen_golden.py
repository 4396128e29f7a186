"""CPU: the oracle against the REAL reference pinned on the reference's own fixtures (tests/golden/photogen_*, pins.npz;
written by oracle/make_golden_photogen.py from /root/reference): all 20 bundled scenes x all 90 edits through
set_foreground + transform_depth_pc, points_to_depth called directly at 1024^2 (config 5 A/B/C), and the guidance weight
schedule looked up by the reference's StepGuidanceWeightSchedule class."""
import os

import numpy as np
import pytest

from oracle import dh_oracle as O
from helpers import sha, f32_translation, load_photogen, photogen_filled_bg
from conftest import GOLDEN_DIR

META, GET = load_photogen()
SCENES = sorted(META)
K_NP = O.get_depth_intrinsics()


def test_pack_is_complete():
    assert len(SCENES) == 20 and sum(len(META[s]["edits"]) for s in SCENES) == 90


@pytest.mark.parametrize("scene", SCENES)
def test_oracle_on_bundled_scene(scene):
    """set_foreground (diffusion_handles.py:105-108) and every edit of transforms.json: bit-exact with the reference."""
    depth, bg, mask = GET(scene)
    m = META[scene]
    assert int(mask.sum()) == m["n_fg"]
    bg2 = photogen_filled_bg(scene, GET)
    assert bg2.dtype == np.float32 and sha(bg2) == m["sha_set_foreground"]
    rows, vals = GET.rows(scene)
    assert np.array_equal(bg2[rows], vals)
    for name, e in m["edits"].items():
        o = O.transform_depth_pc(depth, bg2, mask, K_NP, e["rotation_angle"], e["rotation_axis"], f32_translation(e["translation"]))
        assert o["correspondences"].shape[0] == e["n_corr"], name
        assert sha(o["correspondences"]) == e["sha_corr"], name                  # values AND order
        assert sha(o["disparity"]) == e["sha_disparity"], name                   # Poisson-filled, fp32


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_points_to_depth_1024_pin(case):
    """depth_transform.py:643-747 called directly at 1024^2 on the config-5 point sets (the reference's loop, ~2 s each)."""
    from oracle.make_golden_photogen import config5_points
    pins = np.load(os.path.join(GOLDEN_DIR, "pins.npz"))
    o, S = config5_points(case)

    def pin(k):
        return bytes(pins[f"p2d1024_{case}/sha_{k}"]).hex()
    assert sha(o["points"]) == pin("points")
    pm = np.arange(len(o["points"])) >= S * S
    dm, mk, tx, ty, vis, _ = O.points_to_depth(o["points"], K_NP, (S, S), pm)
    assert sha(dm) == pin("depth_map") and sha(np.packbits(mk)) == pin("depth_mask")
    assert sha(np.packbits(vis)) == pin("visible") and int(vis.sum()) == int(pins[f"p2d1024_{case}/n_visible"])
    assert sha(np.asarray(tx, np.int64)) == pin("tx") and sha(np.asarray(ty, np.int64)) == pin("ty")


@pytest.mark.parametrize("kind", ["constant", "linear", "quadratic"])
@pytest.mark.parametrize("w", [(1.5, 1.25), (1.0, 2.0)])
def test_weight_schedule_pin(kind, w):
    """guided_stable_diffuser.py:336-373 + the reference's StepGuidanceWeightSchedule (:622-665): the product's schedule and the
    oracle's give the reference's weights for every (denoising step, optimisation step)."""
    from diffusionhandles_b200.guided_stable_diffuser import make_guidance_weight_schedule
    table = np.load(os.path.join(GOLDEN_DIR, "pins.npz"))[f"schedule/{kind}_{w[0]}_{w[1]}"]
    ours = make_guidance_weight_schedule(w[0], w[1], 38, kind)
    orc = O.guidance_weight_schedule(w[0], w[1], 38, kind)
    for t in range(table.shape[0]):
        for it in range(table.shape[1]):
            for sched in (ours, orc):
                fg, bg = sched(t, it)
                assert np.array_equal(np.asarray(fg, np.float64), table[t, it, 0]) and np.array_equal(np.asarray(bg, np.float64), table[t, it, 1])
    with pytest.raises(ValueError):
        ours(-1, 0)
