"""Drop-in sampler: the reference's ``SpacedDiffusion`` / ``GaussianDiffusion`` surface that
``val_TDiff`` uses (gaussian_diffusion.py:494-645, respace.py:63-123), driving the fused CUDA loop.

Schedule tables are built in float64 numpy exactly like gaussian_diffusion.py:171-212; with
eta = 0 and START_X prediction the DDIM update collapses to x_{t-1} = a_t*pred + b_t*x_t, which
is the epilogue of the final-layer kernel.  The hypothesis mean + clamp (gaussian_diffusion.py:639-640)
is one more tiny kernel.  Nothing is computed with torch ops on the hot path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .model import DiT, remap_t


def get_named_beta_schedule(schedule_name: str, num_diffusion_timesteps: int) -> np.ndarray:
    """gaussian_diffusion.py:31-75."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_diffusion_timesteps
        return np.array([min(1 - ab((i + 1) / n) / ab(i / n), 0.999) for i in range(n)], dtype=np.float64)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps: int, section_counts):
    """respace.py:7-60 (integer section counts and 'ddimN')."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {desired} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start, all_steps = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            all_steps.append(start + round(cur))
            cur += frac
        start += size
    return set(all_steps)


class SpacedDiffusion:
    """Schedule + DDIM loop.  Only what inference needs (START_X, eta=0, rescaled timesteps)."""

    def __init__(self, use_timesteps, betas, rescale_timesteps: bool = True, predict_xstart: bool = True):
        if not predict_xstart:
            raise NotImplementedError("dvd_b200 supports predict_xstart=True (ModelMeanType.START_X) only")
        base_betas = np.array(betas, dtype=np.float64)
        self.original_num_steps = len(base_betas)
        self.use_timesteps = set(use_timesteps)
        base_acp = np.cumprod(1.0 - base_betas)
        last, new_betas, self.timestep_map = 1.0, [], []
        for i, acp in enumerate(base_acp):                                    # respace.py:77-84
            if i in self.use_timesteps:
                new_betas.append(1 - acp / last)
                last = acp
                self.timestep_map.append(i)
        self.betas = np.array(new_betas, dtype=np.float64)
        assert (self.betas > 0).all() and (self.betas <= 1).all()
        self.num_timesteps = len(self.betas)
        self.rescale_timesteps = rescale_timesteps
        self.alphas_cumprod = np.cumprod(1.0 - self.betas)                     # gaussian_diffusion.py:189-198
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.settings = None                                                   # val_TDiff.py:52 sets it

    # ---- per-step scalars
    def scaled_t(self, i: int) -> float:
        """respace.py:118-122: map_tensor[ts].float() * (1000 / original_num_steps), fp32 arithmetic."""
        t = np.float32(self.timestep_map[i])
        return float(t * np.float32(1000.0 / self.original_num_steps)) if self.rescale_timesteps else float(t)

    def ddim_ab(self, i: int):
        """eta = 0 collapse of gaussian_diffusion.py:470-489."""
        sp = math.sqrt(1.0 - self.alphas_cumprod_prev[i])
        a = math.sqrt(self.alphas_cumprod_prev[i]) - sp / self.sqrt_recipm1_alphas_cumprod[i]
        b = sp * self.sqrt_recip_alphas_cumprod[i] / self.sqrt_recipm1_alphas_cumprod[i]
        return a, b

    def _plan(self):
        idx = list(range(self.num_timesteps))[::-1]                            # gaussian_diffusion.py:564
        t_scaled = [self.scaled_t(i) for i in idx]
        ab = [self.ddim_ab(i) for i in idx]
        return t_scaled, [remap_t(t) for t in t_scaled], [a for a, _ in ab], [b for _, b in ab]

    # ---- the call the evaluation driver makes (evaluation.py:121-135)
    @torch.no_grad()
    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None, device=None,
                         progress=False, eta=0.0, sampling_kwargs=None, logger=None, n_batch=1, time_variant=False, pyramid=None,
                         x_T=None):
        """Returns ``(sample[docs,2,64,64], final_dict)`` like gaussian_diffusion.py:494-535.

        Extensions over the reference: ``shape[0]`` may be a batch of documents (the reference is
        limited to one); ``x_T`` supplies the initial noise explicitly ([docs*n_batch,2,64,64],
        document-major).  When ``x_T`` is None the noise is drawn with the same two torch.randn calls as
        gaussian_diffusion.py:559-569 so that a seeded run consumes the RNG identically."""
        if not isinstance(model, DiT):
            raise TypeError("dvd_b200 sampler drives dvd_b200.DiT only (no foreign-model path)")
        if eta != 0.0:
            raise NotImplementedError("eta != 0 (stochastic DDIM) is not part of the val_TDiff path")
        if clip_denoised or denoised_fn is not None:
            raise NotImplementedError("clip_denoised / denoised_fn are not used by val_TDiff (local.py:73)")
        kw = dict(model_kwargs or {})
        if kw.get("iter") is not True or time_variant is not True:
            raise NotImplementedError("dvd_b200 implements iter=True, time_variant=True (local.py:27-29)")
        if kw.get("src_feat") is not None:
            raise NotImplementedError("src_feat path (train_VGG=False) is dead in the default config")
        dev = model.device
        if dev.type != "cuda":
            raise RuntimeError("dvd_b200 sampler has no CPU path: move the model to a CUDA device")
        docs = int(shape[0])
        assert tuple(shape[1:]) == (2, 64, 64), shape
        with torch.cuda.device(dev):
            if x_T is None:
                _ = noise if noise is not None else torch.randn(*shape, device=dev)        # drawn, then discarded (GD:559-562)
                x_T = torch.randn((docs * n_batch, *shape[1:]), device=dev)                # GD:569
            f = lambda v: v.to(device=dev, dtype=torch.float32).contiguous()
            x_T = f(x_T)
            assert tuple(x_T.shape) == (docs * n_batch, 2, 64, 64)
            eng = model.engine(docs, n_batch)
            eng.static_forward(f(kw["y512"]), f(kw["mask_cat"]), f(kw["mask_y512"]), f(kw["line_msk"]))
            t_scaled, t_emb, a, b = self._plan()
            tables = eng.tables(t_emb)                    # cached on the packed weights (dropped by load_state_dict / .to())
            init_feat0 = kw.get("init_feat")
            if init_feat0 is not None and (t_scaled[0] > 600 or not bool(torch.any(init_feat0 != 0))):
                init_feat0 = None
            out = torch.empty((docs, 2, 64, 64), dtype=torch.float32, device=dev)
            eng.sample(x_T, f(kw["init_flow"]), tables, t_scaled, a, b, None if init_feat0 is None else f(init_feat0), out)
            feat = eng.feat_nhwc().permute(0, 3, 1, 2).clone()    # detached from the workspace the next call overwrites
        final = {"sample": out, "pred_xstart": out, "feat_dict": feat}
        return out, final


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear", use_kl=False,
                              predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False,
                              timestep_respacing=""):
    """script_util.py:206-244."""
    if learn_sigma:
        raise NotImplementedError("learn_sigma=True is not used by val_TDiff (local.py:60)")
    betas = get_named_beta_schedule(noise_schedule, steps)
    if not timestep_respacing:
        timestep_respacing = [steps]
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
                           rescale_timesteps=rescale_timesteps, predict_xstart=predict_xstart)
