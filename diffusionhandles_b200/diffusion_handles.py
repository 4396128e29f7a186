"""``DiffusionHandles`` facade with the reference's method names and signatures (diffusion_handles.py:13-166).

The geometry half of ``transform_foreground`` (:143-149) runs on the sm_100a kernels of this package.  The
diffusion half (U-Net, VAE, text encoder, scheduler, null-text inversion) is stock PyTorch/diffusers code that this
build deliberately does not re-implement (north star): it is injected as ``diffuser`` / ``inverter`` objects that
expose the reference's ``initial_inference`` / ``guided_inference`` / ``invert`` methods.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from .depth_transform import normalize_depth, transform_depth
from .guided_stable_diffuser import GuidedStableDiffuser

DEFAULT_CONF = SimpleNamespace(   # diffhandles/config/default.yaml
    guided_diffuser=SimpleNamespace(bg_weight=1.25, fg_weight=1.5, fg_patch_size=1, bg_patch_size=1, use_depth=True,
                                    save_denoising_steps=False, bg_loss_type='global_avg', num_timesteps=50, num_optsteps=3,
                                    guidance_max_step=38, guidance_schedule_type='constant', bg_erosion=0, seed=2773),
    depth_transform_mode='pc')


class DiffusionHandles:
    def __init__(self, conf=None, diffuser=None, inverter=None):
        self.conf = DEFAULT_CONF if conf is None else conf
        self.diffuser = GuidedStableDiffuser(conf=self.conf.guided_diffuser) if diffuser is None else diffuser
        self.inverter = inverter
        self.device = torch.device('cpu')

    def to(self, device: torch.device = None):
        self.diffuser.to(device=device)
        if self.inverter is not None:
            self.inverter.to(device=device)
        self.device = device
        return self

    def _need(self, obj, method):
        if obj is None or not hasattr(obj, method):
            raise NotImplementedError(
                f"{method} needs the stock-PyTorch diffusion model (U-Net / VAE / scheduler), which is outside this build; "
                "pass a diffuser / inverter object that implements it")
        return getattr(obj, method)

    def invert_input_image(self, img: torch.Tensor, depth: torch.Tensor, prompt: str):
        disparity = normalize_depth(1.0 / depth)
        _, init_noise, null_text_emb = self._need(self.inverter, "invert")(
            target_img=img, depth=disparity, prompt=prompt, num_inner_steps=5, verbose=True)
        return null_text_emb, init_noise

    def generate_input_image(self, depth: torch.Tensor, prompt: str, null_text_emb: torch.Tensor = None,
                             init_noise: torch.Tensor = None):
        disparity = normalize_depth(1.0 / depth)
        with torch.no_grad():
            activations, latent_image, null_text_emb, init_noise = self._need(self.diffuser, "initial_inference")(
                init_latents=init_noise, depth=disparity, uncond_embeddings=null_text_emb, prompt=prompt)
        return null_text_emb, init_noise, activations, latent_image

    def set_foreground(self, depth: torch.Tensor, fg_mask: torch.Tensor, bg_depth: torch.Tensor) -> torch.Tensor:
        """diffusion_handles.py:88-110: background depth = input depth with the (15 px dilated) foreground hole filled by
        a Poisson problem driven by the Laplacian of ``bg_depth``.  Entirely on the device: the dilation is one pass of the
        bit-packed morphology kernel with the 31 x 31 diamond that 15 iterations of SciPy's cross element amount to."""
        from . import _native as N
        from .depth_transform import _pack_mask, _poisson_device
        lib = N.load()
        dev = depth.device
        H, W = depth.shape[-2:]
        d = depth.reshape(1, H, W).to(torch.float32).contiguous()
        b = bg_depth.reshape(1, H, W).to(device=dev, dtype=torch.float32).contiguous()
        bits = _pack_mask((fg_mask.reshape(1, H, W).to(dev) != 0).to(torch.float32).contiguous())
        r = 15
        rows = N.u32_array([((1 << (2 * (r - abs(i - r)) + 1)) - 1) << abs(i - r) for i in range(2 * r + 1)])
        dil = torch.empty_like(bits)
        N.check(lib.dh_morph_pass(N.ptr(bits), N.ptr(dil), 1, H, W, rows, 2 * r + 1, 2 * r + 1, 1, N.stream_handle(dev)), "dh_morph_pass")
        # (the reference's set_foreground is synchronous CPU code: the convergence check's small read-back costs nothing extra)
        return _poisson_device(d, dil, lap_source=b, check=True, what="set_foreground")[None]

    def edit_geometry(self, depth: torch.Tensor, fg_mask: torch.Tensor, bg_depth: torch.Tensor, rot_angle: float = None,
                      rot_axis: torch.Tensor = None, translation: torch.Tensor = None, use_input_depth_normalization=False):
        """The geometry half of an edit on the sm_100a kernels: edited disparity (1,1,H,W) and correspondences (N,4) int64."""
        with torch.no_grad():
            return transform_depth(depth=depth, bg_depth=bg_depth, fg_mask=fg_mask,
                                   intrinsics=self.diffuser.get_depth_intrinsics(device=depth.device), rot_angle=rot_angle,
                                   rot_axis=rot_axis, translation=translation,
                                   use_input_depth_normalization=use_input_depth_normalization,
                                   depth_transform_mode=self.conf.depth_transform_mode)

    def transform_foreground(self, depth: torch.Tensor, prompt: str, fg_mask: torch.Tensor, bg_depth: torch.Tensor,
                             null_text_emb: torch.Tensor, init_noise: torch.Tensor, activations: list,
                             rot_angle: float = None, rot_axis: torch.Tensor = None, translation: torch.Tensor = None,
                             fg_weight: float = None, bg_weight: float = None, use_input_depth_normalization=False):
        """diffusion_handles.py:110-166: geometry on this package's kernels, then the injected diffuser's guided denoising.
        Returns (edited_img, edited_disparity) plus the saved denoising steps when the configuration asks for them."""
        guided = self._need(self.diffuser, "guided_inference")
        disparity, correspondences = self.edit_geometry(depth, fg_mask, bg_depth, rot_angle, rot_axis, translation,
                                                        use_input_depth_normalization)
        keep_steps = bool(self.conf.guided_diffuser.save_denoising_steps)
        with torch.no_grad():
            out = guided(latents=init_noise, depth=disparity, uncond_embeddings=null_text_emb, prompt=prompt,
                         activations_orig=activations, correspondences=correspondences, fg_weight=fg_weight,
                         bg_weight=bg_weight, save_denoising_steps=keep_steps)
        if keep_steps:
            image, steps = out
            return image, disparity, steps
        return out, disparity
