"""CPU: host logic of SURVEY.md 8(f) ranks 3-4 - the DDIM coefficient table (``guided_loop.DDIMSchedule``), the oracle's
restatement of the latent update / CFG combine / DDIM update, and the activation recorder of the generation pass.

The scheduler arithmetic belongs to diffusers 0.23.* (pyproject.toml:30 of the reference), which is not installed: parity is
unpinned against it.  What is checked here: the table against Stable Diffusion's published endpoints, the product table
against the oracle's independent NumPy restatement (bit-exact alphas, coefficients within one ulp: torch's CPU sqrt is not
correctly rounded), and the oracle's elementwise arithmetic against the op sequence DDIMScheduler.step runs, in torch."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import dh_oracle as O
from diffusionhandles_b200.guided_loop import DDIMSchedule
from diffusionhandles_b200.identity import ActivationRecorder, InputImageIdentity


def _coeffs(c):
    return np.array([c.sqrt_beta_t, c.sqrt_alpha_t, c.sqrt_alpha_prev, c.sqrt_beta_prev], dtype=np.float32)


def test_alpha_table_known_answers_and_oracle():
    s = DDIMSchedule()                              # the reference's constructor arguments, guided_stable_diffuser.py:31-32
    a = s.alphas_cumprod.numpy()
    assert a.shape == (1000,) and a.dtype == np.float32
    # Stable Diffusion's scaled-linear schedule: alphas_cumprod[0] = 0.99915, alphas_cumprod[999] = 0.00466 (published values)
    assert abs(float(a[0]) - 0.99915) < 5e-7 and abs(float(a[-1]) - 0.0046601) < 5e-8
    assert np.all(np.diff(a) < 0)
    assert np.array_equal(a, O.ddim_alphas_cumprod())                      # independent NumPy restatement, bit-exact
    assert float(s.final_alpha_cumprod) == float(a[0])                     # set_alpha_to_one=False
    assert float(DDIMSchedule(set_alpha_to_one=True).final_alpha_cumprod) == 1.0


def test_timesteps_and_coefficients():
    s = DDIMSchedule()
    with pytest.raises(RuntimeError):
        s.coefficients(980)
    ts = s.set_timesteps(50)
    assert ts.dtype == np.int64 and ts.tolist() == list(range(980, -1, -20))
    assert np.array_equal(ts, O.ddim_timesteps(50))
    a = O.ddim_alphas_cumprod()
    for t in ts:
        mine, ref = _coeffs(s.coefficients(int(t))), np.array(O.ddim_coefficients(int(t), a, 50), dtype=np.float32)
        assert np.abs(mine.view(np.int32) - ref.view(np.int32)).max() <= 1, t          # torch CPU sqrt vs correctly rounded sqrt
    c0 = s.coefficients(0)                          # last step: the previous alpha is final_alpha_cumprod = alphas_cumprod[0]
    assert c0.sqrt_alpha_prev == c0.sqrt_alpha_t and c0.sqrt_beta_prev == c0.sqrt_beta_t
    assert s.coefficients(980) is s.coefficients(980)                      # cached per timestep
    assert s.coefficients(980, divide_by_reciprocal=False).divide_by_reciprocal == 0
    with pytest.raises(IndexError):
        s.coefficients(1000)
    with pytest.raises(ValueError):
        s.set_timesteps(1001)


def _diffusers_step(model_output, t, sample, sched: DDIMSchedule):
    """DDIMScheduler.step of diffusers 0.23 (epsilon prediction, eta = 0) written with the same torch expressions."""
    prev = t - sched.num_train_timesteps // sched.num_inference_steps
    alpha_prod_t = sched.alphas_cumprod[t]
    alpha_prod_t_prev = sched.alphas_cumprod[prev] if prev >= 0 else sched.final_alpha_cumprod
    beta_prod_t = 1 - alpha_prod_t
    pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
    variance = ((1 - alpha_prod_t_prev) / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)
    std_dev_t = 0.0 * variance ** 0.5
    pred_sample_direction = (1 - alpha_prod_t_prev - std_dev_t ** 2) ** 0.5 * model_output
    return alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction


def test_oracle_step_is_the_scheduler_op_sequence():
    s = DDIMSchedule()
    s.set_timesteps(50)
    a = O.ddim_alphas_cumprod()
    g = torch.Generator().manual_seed(0)
    u, tx, x = (torch.randn(1, 4, 64, 64, generator=g) for _ in range(3))
    for t in (980, 640, 500, 20, 0):
        eps = u + 7.5 * (tx - u)                                            # guided_stable_diffuser.py:470-471
        assert np.array_equal(O.cfg_combine(u.numpy(), tx.numpy()), eps.numpy())
        ref = _diffusers_step(eps, t, x, s).numpy()
        same = O.ddim_step(eps.numpy(), t, x.numpy(), a, 50, coefficients=_coeffs(s.coefficients(t)))
        assert np.array_equal(same, ref), t                                 # elementwise arithmetic: bit-exact
        # a last-ulp change of a coefficient or of the division is amplified by 1 / sqrt(a_t) (13 at t = 980) and the two terms
        # of the update cancel: 1e-5 is the fp32 tolerance of the update, not 1e-6
        own = O.ddim_step(eps.numpy(), t, x.numpy(), a, 50)                 # oracle's own (correctly rounded) coefficients
        assert np.allclose(own, ref, rtol=1e-5, atol=1e-5), t
        rec = O.ddim_step(eps.numpy(), t, x.numpy(), a, 50, reciprocal_division=True)
        assert np.allclose(rec, ref, rtol=1e-5, atol=1e-5), t
    lat, grad = torch.randn(1, 4, 64, 64, generator=g), torch.randn(1, 4, 64, 64, generator=g)
    assert np.array_equal(O.latent_update(lat.numpy(), grad.numpy()), (lat - grad * 0.1).numpy())      # :434


def test_from_scheduler_duck_typing():
    base = DDIMSchedule()
    cfg = dict(prediction_type="epsilon", clip_sample=False, thresholding=False, timestep_spacing="leading", steps_offset=0)
    fake = SimpleNamespace(alphas_cumprod=base.alphas_cumprod.clone(), final_alpha_cumprod=base.alphas_cumprod[0].clone(),
                           config=SimpleNamespace(**cfg), num_inference_steps=50)
    s = DDIMSchedule.from_scheduler(fake)
    assert s is not None and s.num_inference_steps == 50
    assert np.array_equal(_coeffs(s.coefficients(500)), _coeffs((base.set_timesteps(50), base.coefficients(500))[1]))
    fake.config = cfg                                                       # a plain dict config works too
    assert DDIMSchedule.from_scheduler(fake) is not None
    for key, bad in (("prediction_type", "v_prediction"), ("clip_sample", True), ("thresholding", True), ("timestep_spacing", "trailing")):
        fake.config = SimpleNamespace(**{**cfg, key: bad})
        assert DDIMSchedule.from_scheduler(fake) is None, key
    assert DDIMSchedule.from_scheduler(SimpleNamespace(step=lambda *a: None)) is None      # not a DDIM scheduler at all


def test_fused_steps_have_no_cpu_path():
    from diffusionhandles_b200 import _native as N
    from diffusionhandles_b200.guided_loop import cfg_ddim_step, latent_step
    s = DDIMSchedule()
    s.set_timesteps(50)
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(N.NativeLibraryError):
        latent_step(x, x)
    with pytest.raises(N.NativeLibraryError):
        cfg_ddim_step(x, x, x, s.coefficients(980))


def test_activation_recorder():
    T, shapes = 4, [(6, 8, 8), (5, 16, 16), (3, 16, 16)]
    g = torch.Generator().manual_seed(1)
    steps = [[torch.randn(1, *sh, generator=g) for sh in shapes] for _ in range(T)]
    rec = ActivationRecorder(T)
    with pytest.raises(RuntimeError):
        rec.stacks()
    for t in (2, 0, 3):                                                     # any order
        rec.record(t, steps[t])
    with pytest.raises(RuntimeError, match=r"\[1\]"):
        rec.stacks()
    rec.record(1, [a[0] for a in steps[1]])                                 # (C,h,w) maps without the batch dimension
    stacks = rec.stacks()
    # what the reference builds: torch.stack of the per-step activations[0] (guided_stable_diffuser.py:236-239, :269-272)
    for l in range(3):
        assert torch.equal(stacks[l], torch.stack([steps[t][l][0] for t in range(T)], dim=0))
        assert stacks[l].is_contiguous() and stacks[l].shape == (T, *shapes[l])
    with pytest.raises(IndexError):
        rec.record(T, steps[0])
    with pytest.raises(ValueError):
        rec.record(0, steps[0][:2])
    with pytest.raises(ValueError):
        rec.record(0, [steps[0][1], steps[0][1], steps[0][2]])
    ident = rec.identity(torch.zeros(T, 1, 2, 2), torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8))
    assert isinstance(ident, InputImageIdentity) and ident.recorded(2)[1].shape == shapes[1]
    assert ident.nbytes() == 4 * T * sum(c * h * w for c, h, w in shapes)


def test_identity_npz_layout_and_cpu_load(tmp_path):
    """The uncompressed .npz the reference's drivers write (test/test_diffusion_handles.py:88-114): the loader finds every
    array's raw bytes inside the archive (what it memory-maps and streams to the device)."""
    from diffusionhandles_b200.identity import KEYS, load_identity, npz_member_layout, save_identity
    g = torch.Generator().manual_seed(2)
    ident = InputImageIdentity(null_text_emb=torch.randn(3, 1, 7, 8, generator=g), init_noise=torch.randn(1, 4, 8, 8, generator=g),
                               activations=[torch.randn(3, 6, 4, 4, generator=g), torch.randn(3, 5, 8, 8, generator=g),
                                            torch.randn(3, 2, 8, 8, generator=g)], latent_image=torch.randn(1, 4, 8, 8, generator=g))
    path = str(tmp_path / "input_image_identity.npz")
    save_identity(path, ident)
    layout = npz_member_layout(path)
    assert layout is not None and sorted(layout) == sorted(KEYS)
    with np.load(path) as z:
        for k, (offset, shape, dtype) in layout.items():
            assert shape == z[k].shape and dtype == z[k].dtype
            assert np.array_equal(np.memmap(path, dtype=dtype, mode="r", offset=offset, shape=shape), z[k]), k
    back = load_identity(path, "cpu")
    assert all(torch.equal(a, b) for a, b in zip(back.activations, ident.activations)) and torch.equal(back.init_noise, ident.init_noise)
    packed = str(tmp_path / "compressed.npz")
    np.savez_compressed(packed, **{k: np.zeros(3, dtype=np.float32) for k in KEYS})
    assert npz_member_layout(packed) is None                               # compressed members: the np.load path
    assert load_identity(packed, "cpu").latent_image.shape == (3,)
    np.savez(str(tmp_path / "other.npz"), x=np.zeros(3))
    with pytest.raises(KeyError):
        load_identity(str(tmp_path / "other.npz"), "cpu")


def test_oracle_step_equals_the_published_ddim_update():
    """DDIM (Song et al. 2021), eq. 12 with sigma_t = 0, rearranged: x_{t-1} = sqrt(a_prev / a_t) x_t +
    (sqrt(1 - a_prev) - sqrt(a_prev (1 - a_t) / a_t)) eps.  The oracle's fp32 step must agree with this closed form evaluated in
    fp64 - an anchor that does not go through the scheduler's intermediate x0 at all."""
    a = O.ddim_alphas_cumprod().astype(np.float64)
    rng = np.random.default_rng(3)
    eps, x = rng.standard_normal((4, 64, 64)), rng.standard_normal((4, 64, 64))
    for t in O.ddim_timesteps(50):
        t = int(t)
        a_t, a_p = a[t], (a[t - 20] if t >= 20 else a[0])
        closed = np.sqrt(a_p / a_t) * x + (np.sqrt(1 - a_p) - np.sqrt(a_p * (1 - a_t) / a_t)) * eps
        got = O.ddim_step(eps.astype(np.float32), t, x.astype(np.float32), O.ddim_alphas_cumprod(), 50)
        assert np.abs(got - closed).max() <= 2e-5 * max(1.0, np.abs(closed).max()), t
    # chaining all 50 updates with eps = 0 scales the sample by sqrt(a_final / a_980): the telescoping product of the first factors
    y = x.astype(np.float32)
    for t in O.ddim_timesteps(50):
        y = O.ddim_step(np.zeros_like(y), int(t), y, O.ddim_alphas_cumprod(), 50)
    assert np.allclose(y, np.sqrt(a[0] / a[980]) * x, rtol=1e-4, atol=1e-5)
