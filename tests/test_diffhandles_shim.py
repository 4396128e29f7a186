"""CPU: the reference's import names work without any install call (diffhandles/__init__.py:1 and the hot-path modules)."""
import importlib


def test_import_names_resolve_to_the_b200_modules():
    import diffhandles
    import diffusionhandles_b200 as impl
    assert diffhandles.DiffusionHandles is impl.diffusion_handles.DiffusionHandles
    for name in ("depth_transform", "losses", "renderer", "pytorch3d_renderer", "mesh", "utils", "guided_stable_diffuser",
                 "diffusion_handles"):
        m = importlib.import_module(f"diffhandles.{name}")
        assert m is importlib.import_module(f"diffusionhandles_b200.{name}")
    from diffhandles.depth_transform import transform_depth, transform_depth_pc, points_to_depth, depth_to_world_coords  # noqa: F401
    from diffhandles.losses import compute_foreground_loss, compute_background_loss  # noqa: F401
    from diffhandles.guided_stable_diffuser import GuidedStableDiffuser, StepGuidanceWeightSchedule  # noqa: F401
    from diffhandles.utils import pack_correspondences, unpack_correspondences, solve_laplacian_depth  # noqa: F401
    from diffhandles.renderer import Camera, Renderer, RendererArgs  # noqa: F401
    import inspect
    sig = inspect.signature(GuidedStableDiffuser.guided_inference)
    assert list(sig.parameters)[1:] == ["latents", "depth", "uncond_embeddings", "prompt", "activations_orig", "correspondences",
                                        "fg_weight", "bg_weight", "save_denoising_steps"]      # guided_stable_diffuser.py:291-294
