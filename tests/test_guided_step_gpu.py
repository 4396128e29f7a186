"""GPU (``-m gpu``): SURVEY.md 8(f) rank 4 - the fused elementwise steps around the U-Net (``dh_latent_step``,
``dh_cfg_ddim_step``, through ctypes -> C ABI) against the oracle and against the sequence of torch ops the reference runs
(guided_stable_diffuser.py:434, :470-474 with diffusers' DDIMScheduler.step), and the loops that use them
(``guided_inference`` with a DDIM scheduler injected, ``initial_inference`` with the activation recorder)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import dh_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _coeffs(c):
    return np.array([c.sqrt_beta_t, c.sqrt_alpha_t, c.sqrt_alpha_prev, c.sqrt_beta_prev], dtype=np.float32)


class TinyDDIM:
    """DDIMScheduler of diffusers 0.23 as the reference constructs it (guided_stable_diffuser.py:31-32), restated with the same
    torch expressions: CPU fp32 alphas_cumprod, 0-d coefficient tensors multiplied into device tensors."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.config = SimpleNamespace(num_train_timesteps=1000, prediction_type="epsilon", clip_sample=False, thresholding=False,
                                      timestep_spacing="leading", steps_offset=0)
        self.num_inference_steps = None

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        ts = (np.arange(0, n) * (1000 // n)).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, x, t):
        return x

    def step(self, model_output, timestep, sample, eta=0.0, generator=None, return_dict=False):
        timestep = int(timestep)
        prev = timestep - 1000 // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        variance = ((1 - alpha_prod_t_prev) / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)
        std_dev_t = eta * variance ** 0.5
        pred_sample_direction = (1 - alpha_prod_t_prev - std_dev_t ** 2) ** 0.5 * model_output
        return (alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction,)


def test_latent_step(dev):
    from diffusionhandles_b200.guided_loop import latent_step
    g = torch.Generator(device=dev).manual_seed(0)
    for n in (1, 3, 4, 7, 1027, 4 * 64 * 64):
        lat, grad = torch.randn(n, generator=g, device=dev), torch.randn(n, generator=g, device=dev)
        out = latent_step(lat, grad, 0.1)
        assert torch.equal(out, lat - grad * 0.1), n                                   # guided_stable_diffuser.py:434, bit-exact
        assert np.array_equal(out.cpu().numpy(), O.latent_update(lat.cpu().numpy(), grad.cpu().numpy(), 0.1)), n
    lat4 = torch.randn((1, 4, 64, 64), generator=g, device=dev)
    grad4 = torch.randn((1, 4, 64, 64), generator=g, device=dev)
    ref = lat4 - grad4 * 0.1
    assert latent_step(lat4, grad4).shape == (1, 4, 64, 64)
    buf = lat4.clone()
    assert latent_step(buf, grad4, out=buf) is buf and torch.equal(buf, ref)           # in place
    big = torch.randn(4 * 64 * 64 + 1, generator=g, device=dev)                         # pointers that are not 16-byte aligned
    un, gr = big[1:], torch.randn(4 * 64 * 64 + 1, generator=g, device=dev)[1:]
    assert torch.equal(latent_step(un, gr, 0.25), un - gr * 0.25)
    with pytest.raises(ValueError):
        latent_step(lat4, grad4[:, :2])
    with pytest.raises(ValueError):
        latent_step(lat4, grad4, out=torch.empty(3, device=dev))
    assert latent_step(lat4[:, :0], grad4[:, :0]).numel() == 0


def test_cfg_ddim_step(dev):
    from diffusionhandles_b200.guided_loop import DDIMSchedule, cfg_ddim_step
    sched, ref_sched = DDIMSchedule(), TinyDDIM()
    sched.set_timesteps(50)
    ref_sched.set_timesteps(50, device=dev)
    a = O.ddim_alphas_cumprod()
    g = torch.Generator(device=dev).manual_seed(1)
    for n in (4 * 64 * 64, 5, 1026):
        u, tx, x = (torch.randn(n, generator=g, device=dev) for _ in range(3))
        for t in (980, 640, 20, 0):
            c = sched.coefficients(t)
            eps_out = torch.empty_like(x)
            out = cfg_ddim_step(u, tx, x, c, eps_out=eps_out)
            # the oracle with the same coefficients and the reciprocal form of the division: bit-exact
            eps_o = O.cfg_combine(u.cpu().numpy(), tx.cpu().numpy(), 7.5)
            assert np.array_equal(eps_out.cpu().numpy(), eps_o), (n, t)
            o = O.ddim_step(eps_o, t, x.cpu().numpy(), a, 50, reciprocal_division=True, coefficients=_coeffs(c))
            assert np.array_equal(out.cpu().numpy(), o), (n, t)
            # the torch ops the reference runs on the device (CFG :470-471, DDIMScheduler.step :474)
            eps_t = u + 7.5 * (tx - u)
            ref = ref_sched.step(eps_t, t, x, eta=0.0)[0]
            assert torch.equal(eps_out, eps_t), (n, t)
            assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5), (n, t)
            # IEEE division variant (what torch computes on the CPU)
            out_div = cfg_ddim_step(u, tx, x, sched.coefficients(t, divide_by_reciprocal=False))
            o_div = O.ddim_step(eps_o, t, x.cpu().numpy(), a, 50, reciprocal_division=False, coefficients=_coeffs(c))
            assert np.array_equal(out_div.cpu().numpy(), o_div), (n, t)
            # bit-identical to the device ops: ATen's CUDA kernels divide by a host scalar as a multiplication by its fp32
            # reciprocal (the default mode); a build that divides instead is matched by the IEEE mode
            assert torch.equal(out, ref) or torch.equal(out_div, ref), (n, t, float((out - ref).abs().max()))
            # the oracle's own (correctly rounded) coefficients: fp32 tolerance of the update (see tests/test_guided_step_host.py)
            assert np.allclose(out.cpu().numpy(), O.ddim_step(eps_o, t, x.cpu().numpy(), a, 50), rtol=1e-5, atol=1e-5)
    # plain DDIM update (no guidance pair), in place, 4-d shapes
    u4, x4 = torch.randn((1, 4, 64, 64), generator=g, device=dev), torch.randn((1, 4, 64, 64), generator=g, device=dev)
    c = sched.coefficients(500)
    plain = cfg_ddim_step(u4, None, x4, c)
    assert plain.shape == x4.shape
    assert np.array_equal(plain.cpu().numpy(), O.ddim_step(u4.cpu().numpy(), 500, x4.cpu().numpy(), a, 50, reciprocal_division=True,
                                                           coefficients=_coeffs(c)))
    buf = x4.clone()
    assert cfg_ddim_step(u4, None, buf, c, out=buf) is buf and torch.equal(buf, plain)
    with pytest.raises(ValueError):
        cfg_ddim_step(u4[:, :2], None, x4, c)


class TinyUNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.config = SimpleNamespace(sample_size=64, in_channels=5)
        self.c0 = torch.nn.Conv2d(5, 6, 3, padding=1, stride=2)
        self.c1 = torch.nn.Conv2d(5, 5, 3, padding=1)
        self.c2 = torch.nn.Conv2d(5, 4, 3, padding=1)
        self.out = torch.nn.Conv2d(5, 4, 3, padding=1)

    def forward(self, x, t, encoder_hidden_states=None, cross_attention_kwargs=None, return_dict=False):
        s = float(t) / 1000.0 + encoder_hidden_states.mean()
        return (self.out(x) * 0.1, None, None, None, torch.tanh(self.c0(x) + s), torch.tanh(self.c1(x) - s), torch.tanh(self.c2(x) * 2 + s))


def _setup(dev, T):
    from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
    torch.manual_seed(0)
    conf = SimpleNamespace(fg_weight=1.5, bg_weight=1.25, fg_patch_size=1, bg_patch_size=1, use_depth=True, bg_loss_type='global_avg',
                           num_timesteps=T, num_optsteps=2, guidance_max_step=2, guidance_schedule_type='constant', bg_erosion=0, seed=7)
    unet = TinyUNet().to(dev)
    for p_ in unet.parameters():
        p_.requires_grad_(False)
    gsd = GuidedStableDiffuser(conf, unet=unet, scheduler=TinyDDIM()).to(dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    latents0 = torch.randn((1, 4, 64, 64), generator=gen, device=dev)
    depth = torch.rand((1, 1, 512, 512), generator=gen, device=dev) + 1.0
    cond = torch.randn((1, 7, 8), generator=gen, device=dev) * 0.1
    uncond = torch.randn((T, 1, 7, 8), generator=gen, device=dev) * 0.1
    return conf, unet, gsd, gen, latents0, depth, cond, uncond


def test_guided_inference_with_a_ddim_scheduler_uses_the_fused_update(dev, golden_pc, monkeypatch):
    """With a DDIM scheduler injected guided_inference replaces `7.5 * (..)` + scheduler.step by ONE launch; the result must be what
    the scheduler's own step() gives (the unfused loop of the same class)."""
    from diffusionhandles_b200 import guided_loop
    meta, g = golden_pc
    corr = torch.from_numpy(g["cfg1/corr"].astype(np.int64))
    T = 3
    conf, unet, gsd, gen, latents0, depth, cond, uncond = _setup(dev, T)
    acts_orig = [torch.randn((T, c, s, s), generator=gen, device=dev) for c, s in ((6, 32), (5, 64), (4, 64))]
    assert gsd._fused_ddim() is not None
    calls = {"ddim": 0, "lat": 0}
    real_ddim, real_lat = guided_loop.cfg_ddim_step, guided_loop.latent_step
    monkeypatch.setattr(guided_loop, "cfg_ddim_step", lambda *a, **k: (calls.__setitem__("ddim", calls["ddim"] + 1), real_ddim(*a, **k))[1])
    monkeypatch.setattr(guided_loop, "latent_step", lambda *a, **k: (calls.__setitem__("lat", calls["lat"] + 1), real_lat(*a, **k))[1])
    fused = gsd.guided_inference(latents0.clone(), depth, uncond, cond, acts_orig, corr)
    assert calls == {"ddim": T, "lat": 2 * 2}               # one fused update per timestep, one latent update per guided iteration
    monkeypatch.setattr(gsd, "_fused_ddim", lambda: None)
    plain = gsd.guided_inference(latents0.clone(), depth, uncond, cond, acts_orig, corr)
    assert calls["ddim"] == T                               # the scheduler's own step() ran this time
    assert fused.shape == plain.shape == (1, 4, 64, 64)
    assert torch.allclose(fused, plain, rtol=1e-4, atol=1e-4), float((fused - plain).abs().max())


def test_initial_inference_records_the_stacks(dev):
    """initial_inference (guided_stable_diffuser.py:155-274): stacks written in place by the recorder == torch.stack of the per-step
    activations; latents == the reference loop written with torch ops."""
    T = 4
    conf, unet, gsd, gen, latents0, depth, cond, uncond = _setup(dev, T)
    acts, lat, unc, init = gsd.initial_inference(latents0.clone(), depth, uncond, cond)
    assert init.shape == latents0.shape and torch.equal(init, latents0) and unc is uncond
    assert [tuple(a.shape) for a in acts] == [(T, 6, 32, 32), (T, 5, 64, 64), (T, 4, 64, 64)]
    sch = TinyDDIM()
    sch.set_timesteps(T, device=dev)
    d64 = gsd.init_depth(depth)
    x = latents0.clone()
    lists = [[], [], []]
    with torch.no_grad():
        for t_idx, t in enumerate(sch.timesteps):
            o = unet(torch.cat([x, d64], dim=1), t, encoder_hidden_states=cond)
            for l in range(3):
                lists[l].append(o[4 + l][0])
            x2 = torch.cat([torch.cat([x] * 2), torch.cat([d64] * 2, dim=0)], dim=1)
            n = unet(x2, t, encoder_hidden_states=torch.cat([uncond[t_idx].expand(*cond.shape), cond]))[0]
            nu, nt = n.chunk(2)
            x = sch.step(nu + 7.5 * (nt - nu), t, x)[0]
    for l in range(3):
        ref = torch.stack(lists[l], dim=0)
        assert torch.allclose(acts[l], ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(lat, x, rtol=1e-4, atol=1e-4)
    from diffusionhandles_b200.guided_stable_diffuser import GuidedStableDiffuser
    with pytest.raises(NotImplementedError):
        GuidedStableDiffuser(conf).initial_inference(latents0, depth, uncond, cond)


def test_whole_edit_through_the_facade(dev, golden_pc, tmp_path):
    """The reference's driver sequence (test/test_diffusion_handles.py:85-200) through the top-level API with tiny stand-in models:
    generate_input_image (recording pass) -> identity .npz round trip -> set_foreground -> transform_foreground (geometry kernels,
    then guided_inference with the fused loss / latent / CFG + DDIM launches).  The correspondences that reach the diffuser are the
    reference's (golden), the edited latents equal the loop written with torch gathers."""
    import torch.nn.functional as F
    from diffusionhandles_b200 import DiffusionHandles
    from diffusionhandles_b200.guided_stable_diffuser import make_guidance_weight_schedule
    from diffusionhandles_b200.identity import load_identity, save_identity
    meta, g = golden_pc
    m = meta["cfg1"]
    T = 3
    conf, unet, gsd, gen, latents0, _, cond, uncond = _setup(dev, T)
    conf.save_denoising_steps = False
    dh = DiffusionHandles(SimpleNamespace(guided_diffuser=conf, depth_transform_mode='pc'), diffuser=gsd).to(dev)
    depth, bg, mask = O.synthetic_scene(**m["scene"])
    td, tb, tm = (torch.from_numpy(a).to(dev)[None, None] for a in (depth, bg, mask))
    # recording pass + identity file
    null_text, init_noise, activations, latent_image = dh.generate_input_image(td, cond, null_text_emb=uncond, init_noise=latents0.clone())
    assert [tuple(a.shape) for a in activations] == [(T, 6, 32, 32), (T, 5, 64, 64), (T, 4, 64, 64)] and torch.equal(init_noise, latents0)
    from diffusionhandles_b200.identity import InputImageIdentity
    path = str(tmp_path / "input_image_identity.npz")
    save_identity(path, InputImageIdentity(null_text_emb=null_text, init_noise=init_noise, activations=activations, latent_image=latent_image))
    back = load_identity(path, dev)
    assert all(torch.equal(a, b) for a, b in zip(back.activations, activations))
    # the edit
    seen = {}
    real = gsd.guided_inference

    def spy(**kw):
        seen["corr"] = kw["correspondences"]
        seen["disparity"] = kw["depth"]
        return real(**kw)
    gsd.guided_inference = spy
    edited, disparity = dh.transform_foreground(td, cond, tm, tb, back.null_text_emb, back.init_noise, back.activations, rot_angle=m["angle"],
                                                rot_axis=torch.tensor(m["axis"]), translation=torch.tensor(m["translation"]))
    assert np.array_equal(seen["corr"].numpy(), g["cfg1/corr"].astype(np.int64))            # the reference's correspondences
    assert disparity.shape == (1, 1, 512, 512) and edited.shape == (1, 4, 64, 64)
    # the same guided denoising written with torch gathers (losses.py:4-84) and the scheduler's own step()
    pc = O.process_correspondences(seen["corr"].numpy(), 512, 0)
    ix = {k: torch.from_numpy(v).to(dev) for k, v in pc.items()}
    sched = make_guidance_weight_schedule(1.5, 1.25, 2, 'constant')
    d64 = gsd.init_depth(seen["disparity"])
    sch = TinyDDIM()
    sch.set_timesteps(T, device=dev)
    lat = latents0.clone()

    def up(a):
        return a if a.shape[-1] == 64 else F.interpolate(a[None], size=(64, 64), mode="bilinear", align_corners=False)[0]
    for t_idx, t in enumerate(sch.timesteps):
        for it in range(2 if t_idx < 2 else 0):
            l_ = lat.detach().requires_grad_(True)
            o_ = unet(torch.cat([l_, d64], dim=1), t, encoder_hidden_states=cond)
            fgw, bgw = sched(t_idx, it)
            loss = 0.0
            for li, a in enumerate((o_[4], o_[5], o_[6])):
                cur, org = up(a[0]), up(activations[li][t_idx])
                fg = (org[:, ix["original_y"], ix["original_x"]] - cur[:, ix["transformed_y"], ix["transformed_x"]]).abs().mean(-1).mean()
                bgl = (org[:, ix["background_y_orig"], ix["background_x_orig"]].mean(-1) -
                       cur[:, ix["background_y_trans"], ix["background_x_trans"]].mean(-1)).abs().mean()
                loss = loss + fgw[li] * fg + bgw[li] * bgl
            lat = l_.detach() - 0.1 * torch.autograd.grad(loss, [l_])[0]
        with torch.no_grad():
            x2 = torch.cat([torch.cat([lat] * 2), torch.cat([d64] * 2, dim=0)], dim=1)
            n = unet(x2, t, encoder_hidden_states=torch.cat([uncond[t_idx].expand(*cond.shape), cond]))[0]
            nu, nt = n.chunk(2)
            lat = sch.step(nu + 7.5 * (nt - nu), t, lat)[0]
    assert torch.allclose(edited, lat, rtol=1e-4, atol=1e-4), float((edited - lat).abs().max())
