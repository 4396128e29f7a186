"""The guided denoising loop around the (stock PyTorch) U-Net - SURVEY.md 8(f) rank 4, mirroring
guided_stable_diffuser.py:377-480: per timestep up to ``num_optsteps`` gradient steps on the latents, driven by the
fused guidance loss (one K4 launch per evaluation instead of six loss calls), then the classifier-free-guidance step.

The U-Net, the scheduler and the text embeddings are injected as callables / tensors; this module contains no model.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .guided_stable_diffuser import make_guidance_weight_schedule
from .losses import guidance_loss, guidance_loss_and_grad


def guided_denoise(latents: torch.Tensor, timesteps: Sequence, unet: Callable, scheduler_step: Callable,
                   activations_orig: List[torch.Tensor], processed_correspondences,
                   fg_weight: float = 1.5, bg_weight: float = 1.25, num_optsteps: int = 3, guidance_max_step: int = 38,
                   guidance_schedule_type: str = "constant", bg_loss_type: str = "global_avg", fg_patch_size: int = 1,
                   bg_patch_size: int = 1, step_size: float = 0.1, scale_model_input: Optional[Callable] = None,
                   cfg_noise: Optional[Callable] = None, skip_zero_weight_layers: bool = False,
                   on_step: Optional[Callable] = None) -> torch.Tensor:
    """latents (1,4,h,w).  ``unet(latents_in, t) -> (noise_pred, [act0, act1, act2])`` with activations (1,C,h,w) that are
    differentiable w.r.t. ``latents_in``; ``scheduler_step(noise_pred, t, latents) -> latents``;
    ``cfg_noise(latents, t, t_idx) -> noise_pred`` runs the classifier-free-guidance forward (defaults to ``unet``).

    guided_stable_diffuser.py:415-434: loss = sum_l fgw[l] * L_fg,l + bgw[l] * L_bg,l; latents -= 0.1 * dloss/dlatents.
    """
    schedule = make_guidance_weight_schedule(fg_weight, bg_weight, guidance_max_step, guidance_schedule_type)
    for t_idx, t in enumerate(timesteps):
        iteration = 0
        while iteration < num_optsteps and t_idx < guidance_max_step:
            with torch.enable_grad():
                lat = latents.detach().requires_grad_(True)
                model_in = scale_model_input(lat, t) if scale_model_input is not None else lat
                _, activations = unet(model_in, t)
                fgw, bgw = schedule(t_idx, iteration)
                acts = [a[0] for a in activations]
                origs = [a[t_idx] for a in activations_orig]
                activations_size = (activations_orig[2][t_idx].shape[-2], activations_orig[2][t_idx].shape[-1])
                if skip_zero_weight_layers:       # a zero-weight layer contributes exactly 0 (the reference still evaluates it)
                    keep = [i for i in range(len(acts)) if fgw[i] != 0.0 or bgw[i] != 0.0]
                    acts, origs = [acts[i] for i in keep], [origs[i] for i in keep]
                    fgw, bgw = [fgw[i] for i in keep], [bgw[i] for i in keep]
                if acts and (bg_loss_type != 'local_avg' or bg_patch_size == fg_patch_size):
                    # value and d(loss)/d(activations) from one fused launch; the U-Net's backward then carries them to the latents
                    # (the chain rule of autograd.grad(loss, latents), guided_stable_diffuser.py:430-434)
                    _, _, act_grads = guidance_loss_and_grad(acts, origs, processed_correspondences, fgw, bgw, bg_loss_type=bg_loss_type,
                                                             activations_size=activations_size, patch_size=fg_patch_size)
                    grad = torch.autograd.grad(acts, [lat], grad_outputs=act_grads)[0]
                    latents = lat.detach() - grad * step_size
                elif acts:
                    loss, _ = guidance_loss(acts, origs, processed_correspondences, fgw, bgw, bg_loss_type=bg_loss_type,
                                            activations_size=activations_size, patch_size=fg_patch_size,
                                            bg_patch_size=bg_patch_size)
                    grad = torch.autograd.grad(loss, [lat])[0]
                    latents = lat.detach() - grad * step_size
            iteration += 1
        if on_step is not None:
            on_step('opt', latents)
        with torch.no_grad():
            noise_pred = cfg_noise(latents, t, t_idx) if cfg_noise is not None else unet(latents, t)[0]
            latents = scheduler_step(noise_pred, t, latents)
            if on_step is not None:
                on_step('post-opt', latents)
    return latents
