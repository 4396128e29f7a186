"""CPU: the NumPy oracle (oracle/dh_oracle.py) against golden vectors produced by the REAL reference
(oracle/make_golden.py).  This is what pins the oracle (parity gate, step 3 of the brief)."""
import numpy as np
import pytest

from oracle import dh_oracle as O
from helpers import sha, f32_translation

K = O.get_depth_intrinsics()


def test_intrinsics_value():
    # guided_stable_diffuser.py:148-153; SURVEY.md 8(a) row 1
    assert float(K[0, 0]) == 1.9209821224212646 and K[2, 2] == 1 and K[0, 2] == 0


def test_linspace_matches_torch_cpu():
    import torch
    for n in (2, 3, 8, 33, 64, 100, 511, 512, 1024, 2048):
        assert np.array_equal(torch.linspace(-1, 1, n).numpy(), O.linspace_f32(-1, 1, n)), n
    for (a, b, n) in ((-0.6, 0.6, 48), (-47 / 79, 47 / 79, 48), (-32 / 69, 32 / 69, 33)):
        assert np.array_equal(torch.linspace(-a if a > 0 else a, b, n).numpy(), O.linspace_f32(a, b, n))


@pytest.mark.parametrize("name", ["cfg1", "neg60", "occl90", "zties45", "xaxis20", "identity", "cfg1_norm",
                                  "zaxis_all_offscreen", "axis_scaled", "smooth25", "smooth_m50"])
def test_pc_transform_against_reference_golden(golden_pc, name):
    meta, g = golden_pc
    m = meta[name]
    depth, bg, mask = O.synthetic_scene(**m["scene"])
    assert sha(depth) + sha(bg) + sha(mask) == m["sha_inputs"], "synthetic scene generator drifted"
    o = O.transform_depth_pc(depth, bg, mask, K, m["angle"], m["axis"], f32_translation(m["translation"]),
                             use_input_depth_normalization=m["norm"])
    assert int(mask.sum()) == m["n_fg"]
    assert sha(o["points"]) == m["sha_points"]                                   # fp64 point array, bit-exact
    assert np.array_equal(o["points"][512 * 512:512 * 512 + 64], g[f"{name}/centroid_points_head"])
    assert sha(o["depth_map"]) == m["sha_depth_map"]                             # fp32 z-buffer
    assert np.array_equal(o["depth_map"][::32], g[f"{name}/depth_map_rows"])
    assert np.array_equal(np.packbits(o["target_mask"]), g[f"{name}/target_mask"])
    assert sha(np.packbits(o["visible"])) == m["sha_visible"]
    assert int(o["visible"].sum()) == int(g[f"{name}/visible_count"])
    assert np.array_equal(o["correspondences"], g[f"{name}/corr"].astype(np.int64))   # incl. order
    assert o["correspondences"].shape[0] == m["n_corr"]
    assert np.array_equal(o["disparity"][::32], g[f"{name}/disparity_rows"])
    assert sha(o["disparity"]) == m["sha_disparity"]


def test_pc_empty_mask_branch(golden_pc):
    meta, _ = golden_pc
    m = meta["empty_mask"]
    depth, bg, _ = O.synthetic_scene(**m["scene"])
    o = O.transform_depth_pc(depth, bg, np.zeros_like(depth), K, 10.0)
    assert o["correspondences"].shape == tuple(m["corr_shape"]) and o["correspondences"].dtype == np.int64
    assert sha(o["disparity"]) == m["sha_disparity"]


def test_zbuffer_closed_form_equals_loop():
    rng = np.random.default_rng(0)
    for trial in range(5):
        n, npix = 5000, 300
        z = np.round(rng.uniform(-1, 3, n), 1)
        pix = rng.integers(0, npix, n)
        assert np.array_equal(O.zbuffer_closed_form(z, pix, npix), O.zbuffer_loop(z, pix, npix))


@pytest.mark.parametrize("tag", ["sq", "wide", "tall"])
def test_unproject_golden(golden_small, tag):
    g = golden_small
    assert np.array_equal(O.depth_to_world_coords(g[f"unproj_{tag}/depth"], K), g[f"unproj_{tag}/points"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_points_to_depth_golden(golden_small, tag):
    g = golden_small
    H, W = g[f"p2d_{tag}/size"]
    for loop in (False, True):
        dm, mk, tx, ty, vis, _ = O.points_to_depth(g[f"p2d_{tag}/points"], K, (int(H), int(W)), g[f"p2d_{tag}/point_mask"], loop=loop)
        assert np.array_equal(dm, g[f"p2d_{tag}/depth_map"])
        assert np.array_equal(mk, g[f"p2d_{tag}/depth_mask"])
        assert np.array_equal(vis, g[f"p2d_{tag}/visible"])
        assert np.array_equal(tx, g[f"p2d_{tag}/tx"]) and np.array_equal(ty, g[f"p2d_{tag}/ty"])


def test_normalize_depth_golden(golden_small):
    g = golden_small
    assert np.array_equal(O.normalize_depth(g["nd/x"][0, 0]), g["nd/y"][0, 0])


def test_ellipse_elements_match_opencv():
    cv2 = pytest.importorskip("cv2")
    for k in range(1, 48):
        assert np.array_equal(O.ellipse_element(k), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))), k


def test_morphology_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for S, p in ((64, 0.3), (128, 0.6), (97, 0.5)):
        m = (rng.random((S, S)) < p)
        m[:5, :7] = True
        m[-3:, -9:] = True
        u8 = m.astype(np.uint8) * 255
        for k in (1, 2, 3, 4, 5, 10, 20):
            el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
            assert np.array_equal(O.morph_dilate(m, el), cv2.dilate(u8, el) == 255), (S, k)
            assert np.array_equal(O.morph_erode(m, el), cv2.erode(u8, el) == 255), (S, k)
            assert np.array_equal(O.morph_erode(O.morph_dilate(m, el), el), cv2.morphologyEx(u8, cv2.MORPH_CLOSE, el) == 255)
            assert np.array_equal(O.morph_dilate(O.morph_erode(m, el), el), cv2.morphologyEx(u8, cv2.MORPH_OPEN, el) == 255)


@pytest.mark.parametrize("tag", ["e0", "e5", "e15", "r1024", "oob"])
def test_process_correspondences_golden(golden_small, tag):
    g = golden_small
    res, er = g[f"pcorr_{tag}/res_er"]
    pc = O.process_correspondences(g[f"pcorr_{tag}/corr"].astype(np.int64), int(res), int(er))
    for k, v in pc.items():
        assert np.array_equal(v, g[f"pcorr_{tag}/{k}"].astype(np.int64)), k


@pytest.mark.parametrize("tag", ["c6h64", "c5h32", "c3h16"])
def test_losses_golden(golden_small, golden_pc, tag):
    g = golden_small
    _, gp = golden_pc
    pc = O.process_correspondences(gp["cfg1/corr"].astype(np.int64), 512, 0)
    cur, orig = g[f"loss_{tag}/cur"], g[f"loss_{tag}/orig"]

    def check(val, grad, key):
        ref_v, ref_g = g[f"loss_{tag}/{key}"], g[f"loss_{tag}/{key}_grad"]
        assert abs(val - ref_v) <= 1e-5 * abs(ref_v), key
        scale = np.abs(ref_g).max()
        assert np.abs(grad - ref_g).max() <= 1e-5 * scale, key
    check(*O.foreground_loss(cur, orig, pc), "fg")
    check(*O.background_loss(cur, orig, pc, loss_type="global_avg"), "bg_global_avg")
    check(*O.background_loss(cur, orig, pc, loss_type="local_avg"), "bg_local_avg")
    with pytest.raises(ValueError):
        O.background_loss(cur, orig, pc, loss_type="nope")


@pytest.mark.parametrize("tag", ["p3c4h64", "p5c3h32", "p2c3h16", "p4c2h64"])
def test_patch_losses_golden(golden_small, golden_pc, tag):
    """patch_size > 1 (losses.py:62-77), odd and even patches, against values/gradients recorded from the reference."""
    g = golden_small
    _, gp = golden_pc
    pc = O.process_correspondences(gp["cfg1/corr"].astype(np.int64), 512, 0)
    cur, orig, patch = g[f"ploss_{tag}/cur"], g[f"ploss_{tag}/orig"], int(g[f"ploss_{tag}/patch"])
    for key, (val, grad) in {"fg": O.foreground_loss(cur, orig, pc, patch=patch),
                             "bg_local_avg": O.background_loss(cur, orig, pc, loss_type="local_avg", patch=patch)}.items():
        ref_v, ref_g = g[f"ploss_{tag}/{key}"], g[f"ploss_{tag}/{key}_grad"]
        assert abs(val - ref_v) <= 1e-5 * abs(ref_v), key
        amb = O.loss_sign_ambiguity(cur, orig, pc, bg_loss_type="local_avg", patch=patch)
        assert amb.mean() < 1e-3
        assert np.where(amb, 0.0, np.abs(grad - ref_g)).max() <= 1e-5 * np.abs(ref_g).max(), key


def test_empty_index_lists_give_nan_loss_and_zero_gradient():
    """torch: the mean over an empty gather is NaN, its backward scatters nothing (verified against the reference)."""
    cur, orig = np.ones((2, 8, 8), np.float32), np.zeros((2, 8, 8), np.float32)
    e = np.zeros(0, np.int64)
    pc = dict(original_x=e, original_y=e, transformed_x=e, transformed_y=e)
    v, gr = O.foreground_loss(cur, orig, pc)
    assert np.isnan(v) and not gr.any()


def test_guidance_weight_schedule_values():
    s = O.guidance_weight_schedule()
    fg, bg = s(0, 0)
    assert fg == [0.0, 0.0, 7.5 * 45.0 * 2.5] and bg == [0.0, 0.0, 1.5 * 37.5 * 1.25]
    fg, bg = s(1, 1)
    assert fg == [0.0, 5.0 * 45.0 * 1.25, 0.0] and bg == [0.0, 1.5 * 37.5 * 2.5, 0.0]
    assert s(38, 0) == ([0.0] * 3, [0.0] * 3) and s(49, 2) == ([0.0] * 3, [0.0] * 3)
    with pytest.raises(ValueError):
        O.guidance_weight_schedule(schedule_type="cubic")


def test_dense_map_and_gather_properties(golden_pc):
    _, gp = golden_pc
    corr = gp["cfg1/corr"].astype(np.int64)
    rng = np.random.default_rng(5)
    for side in (64, 32, 16, 8):
        m = O.dense_source_map(corr, 512, side)
        r = 512 // side
        d = (corr[:, 3] // r) * side + corr[:, 2] // r
        s = (corr[:, 1] // r) * side + corr[:, 0] // r
        assert set(np.nonzero(m >= 0)[0]) == set(d)
        for q in np.unique(d)[:50]:
            assert m[q] == s[np.nonzero(d == q)[0][0]]          # first correspondence in reference order
        A = rng.normal(size=(3, side, side)).astype(np.float32)
        D = O.warp_gather_dense(A, m).reshape(3, -1)
        assert np.array_equal(D[:, m < 0], np.zeros((3, int((m < 0).sum())), np.float32))
        assert np.array_equal(D[:, m >= 0], A.reshape(3, -1)[:, m[m >= 0]])


def test_transform_point_cloud_golden(golden_small):
    g = golden_small
    depth, bg, mask = O.synthetic_scene(S=512, seed=9, radius=70.0)
    pts = O.depth_to_world_coords(depth, K)
    rot, mod = O.transform_point_cloud(pts, np.array([0.0, 1.0, 0.0], np.float32), 33.0, 0.25, -0.5, 0.125, mask)
    assert rot.dtype == np.float64 and rot.shape == (512, 512, 3)
    assert bytes.fromhex(sha(rot)) == g["tpc/sha"].tobytes()
    assert np.array_equal(rot[::64], g["tpc/rows"]) and int(mod.sum()) == int(g["tpc/mod_count"])


@pytest.mark.parametrize("tag", ["s96", "s160"])
def test_solve_laplacian_depth_golden(golden_small, tag):
    import scipy.ndimage
    g = golden_small
    S, it = (int(v) for v in g[f"sld_{tag}/S_it"])
    depth, bg, mask = O.synthetic_scene(S=S, seed=12, radius=S / 6)
    dil = scipy.ndimage.binary_dilation(mask.astype(bool), iterations=it)
    assert np.array_equal(np.packbits(dil), g[f"sld_{tag}/dilated"])
    sol = O.solve_laplacian_depth(depth, bg, dil)
    assert sol.dtype == np.float32 and np.array_equal(sol, g[f"sld_{tag}/solution"])


def test_oracle_rasteriser_sanity():
    """Mesh mode oracle (parity unpinned): a front-facing depth-map mesh covers every pixel, interpolated z is the plane
    depth, a back-facing copy is culled, and exact z ties go to the lower face index."""
    S = 12
    depth = np.full((S, S), 3.0, np.float32)
    v = O.depth_to_world_coords(depth, K).reshape(-1, 3)
    idx = np.arange(S * S).reshape(S, S)
    ul = np.stack([idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[:-1, :-1].ravel()], -1)
    lr = np.stack([idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()], -1)
    faces = np.stack([ul, lr], 1).reshape(-1, 3)
    sx, sy = O.fov_scales(float(K[1, 1]), S, S)
    p2f, z, b = O.rasterize_meshes(v, faces, S, S, sx, sy, 1e-5, True)
    assert (p2f >= 0).all() and np.abs(z - 3.0).max() < 1e-5 and np.abs(b.sum(-1) - 1).max() < 1e-5
    p2f_back, _, _ = O.rasterize_meshes(v, faces[:, ::-1], S, S, sx, sy, 1e-5, True)
    assert (p2f_back < 0).all()
    both = np.concatenate([faces, faces])                      # duplicated mesh: every pixel is an exact z tie
    p2f2, _, _ = O.rasterize_meshes(v, both, S, S, sx, sy, 1e-5, True)
    assert np.array_equal(p2f2, p2f)
    wp = O.interpolate_face_attributes(v, faces, p2f, b)
    assert wp.shape == (S, S, 4) and (wp[..., 3] == 1).all() and np.abs(wp[..., 2] - 3.0).max() < 1e-5
