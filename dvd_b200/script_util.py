"""Factories with the reference's signatures (improved_diffusion/script_util.py:38-90,93-203,257-258)."""
from __future__ import annotations

from .model import DiT_models2
from .sampler import create_gaussian_diffusion


def model_and_diffusion_defaults():
    """script_util.py:11-35."""
    return dict(image_size=256, num_channels=128, num_res_blocks=2, num_heads=4, num_heads_upsample=-1,
                attention_resolutions="16,8", dropout=0.0, learn_sigma=False, sigma_small=False, class_cond=False,
                diffusion_steps=1000, noise_schedule="linear", timestep_respacing="", use_kl=False, predict_xstart=True,
                rescale_timesteps=True, rescale_learned_sigmas=True, use_checkpoint=False, use_scale_shift_norm=True)


def create_model(image_size, num_channels, num_res_blocks, learn_sigma, class_cond, use_checkpoint, attention_resolutions,
                 num_heads, num_heads_upsample, use_scale_shift_norm, dropout, device, train_mode, tv):
    """script_util.py:93-203: only the branch val_TDiff selects (train_mode 'stage_1_dit_cross', :155-162)."""
    if train_mode != "stage_1_dit_cross":
        raise ValueError(f"dvd_b200 implements train_mode='stage_1_dit_cross' only (got {train_mode!r})")
    if image_size != 64:
        raise ValueError(f"unsupported image size: {image_size}")
    return DiT_models2["DiT-S/2"](input_size=64, in_channels=2, tv=tv)


def create_model_and_diffusion(image_size, class_cond, learn_sigma, sigma_small, num_channels, num_res_blocks, num_heads,
                               num_heads_upsample, attention_resolutions, dropout, diffusion_steps, noise_schedule,
                               timestep_respacing, use_kl, predict_xstart, rescale_timesteps, rescale_learned_sigmas,
                               use_checkpoint, use_scale_shift_norm, device, train_mode, tv):
    """script_util.py:38-90 (same positional/keyword surface; called at val_TDiff.py:46-51)."""
    model = create_model(image_size, num_channels, num_res_blocks, learn_sigma=learn_sigma, class_cond=class_cond,
                         use_checkpoint=use_checkpoint, attention_resolutions=attention_resolutions, num_heads=num_heads,
                         num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm, dropout=dropout,
                         device=device, train_mode=train_mode, tv=tv)
    diffusion = create_gaussian_diffusion(steps=diffusion_steps, learn_sigma=learn_sigma, sigma_small=sigma_small,
                                          noise_schedule=noise_schedule, use_kl=use_kl, predict_xstart=predict_xstart,
                                          rescale_timesteps=rescale_timesteps, rescale_learned_sigmas=rescale_learned_sigmas,
                                          timestep_respacing=timestep_respacing)
    return model, diffusion


def args_to_dict(args, keys):
    """script_util.py:257-258."""
    return {k: getattr(args.env, k) for k in keys}
