"""Drop-in sampler: the reference's ``SpacedDiffusion`` / ``GaussianDiffusion`` surface that
``val_TDiff`` uses (gaussian_diffusion.py:494-645, respace.py:63-123), driving the fused CUDA loop.

Schedule tables are built in float64 numpy exactly like gaussian_diffusion.py:171-212; with
eta = 0 and START_X prediction the DDIM update collapses to x_{t-1} = a_t*pred + b_t*x_t, which
is the epilogue of the final-layer kernel.  The hypothesis mean + clamp (gaussian_diffusion.py:639-640)
is one more tiny kernel.  Nothing is computed with torch ops on the hot path.
"""
from __future__ import annotations

import math
import os

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
        document-major).  When ``x_T`` is None the noise is drawn with the same torch.randn calls as
        gaussian_diffusion.py:559-569 (and the S unused per-step draws of :479) so that a seeded run consumes the RNG identically
        and a sequence of documents sees the same noise as under the reference."""
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
                for _ in range(self.num_timesteps):                                        # GD:479 draws randn_like(x) every step and
                    torch.randn_like(x_T)                                                  # multiplies it by sigma = 0: keep the generator in step
            f = lambda v: v.to(device=dev, dtype=torch.float32).contiguous()
            x_T = f(x_T)
            assert tuple(x_T.shape) == (docs * n_batch, 2, 64, 64)
            eng = model.engine(docs, n_batch)
            t_scaled, t_emb, a, b = self._plan()
            tables = eng.tables(t_emb)                    # cached on the packed weights (dropped by load_state_dict / .to())
            init_feat0 = kw.get("init_feat")
            if init_feat0 is not None and (t_scaled[0] > 600 or not bool(torch.any(init_feat0 != 0))):
                init_feat0 = None
            inputs = {"y512": kw["y512"], "mask_cat": kw["mask_cat"], "mask_y512": kw["mask_y512"], "line_msk": kw["line_msk"],
                      "x_T": x_T, "init_flow": kw["init_flow"]}
            if init_feat0 is None and os.environ.get("DVD_NO_GRAPH", "0") != "1":
                # the whole document (static conditioning + S steps + hypothesis mean: ~230 kernels) replays from ONE CUDA graph over
                # static input buffers, like DewarpPipeline: per-call cost = six small device copies + one graph launch
                g = self._graphed(eng, tables, t_scaled, a, b, docs, n_batch, dev)
                out = g.run(inputs)
            else:
                eng.static_forward(f(kw["y512"]), f(kw["mask_cat"]), f(kw["mask_y512"]), f(kw["line_msk"]))
                out = torch.empty((docs, 2, 64, 64), dtype=torch.float32, device=dev)
                eng.sample(x_T, f(kw["init_flow"]), tables, t_scaled, a, b, None if init_feat0 is None else f(init_feat0), out)
            feat = eng.feat_nhwc().permute(0, 3, 1, 2).clone()    # detached from the workspace the next call overwrites
        final = {"sample": out, "pred_xstart": out, "feat_dict": feat}
        return out, final

    @torch.no_grad()
    def ddim_sample_loop_for_training(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None, device=None,
                                      progress=False, eta=0.0, sampling_kwargs=None, logger=None, n_batch=1, time_variant=False, iter=True,
                                      mode="train", timestep=None, pyramid=None, x_T=None):
        """The no-grad roll-out inside ``training_losses_time_variant`` (gaussian_diffusion.py:647-780, called :924-942): DDIM steps
        S-1 .. timestep+1 of ONE sample, with the RAW rescaled timestep embedded (``mode`` is not None, cross_model.py:575), returning
        ``(clamp(pred_xstart), feat)``.  Same kernels as inference; only the step range and the timestep table differ."""
        if not isinstance(model, DiT):
            raise TypeError("dvd_b200 sampler drives dvd_b200.DiT only (no foreign-model path)")
        if n_batch != 1 or eta != 0.0 or clip_denoised or denoised_fn is not None or time_variant is not True or iter is not True or mode is None:
            raise NotImplementedError("roll-out implemented as called at gaussian_diffusion.py:924-942 (n_batch=1, eta=0, tv, iter, mode='train')")
        if timestep is None or not (0 <= int(timestep) < self.num_timesteps - 1):
            raise ValueError("timestep must be in [0, S-2] (the reference rolls out from S-1 down to timestep+1)")
        kw = dict(model_kwargs or {})
        dev = model.device
        if dev.type != "cuda":
            raise RuntimeError("dvd_b200 sampler has no CPU path: move the model to a CUDA device")
        docs = int(shape[0])
        assert docs == 1 and tuple(shape[1:]) == (2, 64, 64), shape
        idx = list(range(int(timestep) + 1, self.num_timesteps))[::-1]                     # GD:723
        with torch.cuda.device(dev):
            if x_T is None:
                _ = noise if noise is not None else torch.randn(*shape, device=dev)        # GD:718-721
                x_T = torch.randn((n_batch, *shape[1:]), device=dev)                        # GD:728
                for _ in idx:
                    torch.randn_like(x_T)                                                  # GD:479 (unused, sigma = 0)
            f = lambda v: v.to(device=dev, dtype=torch.float32).contiguous()
            eng = model.engine(1, 1)
            eng.static_forward(f(kw["y512"]), f(kw["mask_cat"]), f(kw["mask_y512"]), f(kw["line_msk"]))
            t_scaled = [self.scaled_t(i) for i in idx]
            ab = [self.ddim_ab(i) for i in idx]
            tables = eng.tables(t_scaled)                                                  # raw timestep embedding (mode != None)
            init_feat0 = kw.get("init_feat")
            if init_feat0 is not None and (t_scaled[0] > 600 or not bool(torch.any(init_feat0 != 0))):
                init_feat0 = None
            out = torch.empty((1, 2, 64, 64), dtype=torch.float32, device=dev)
            eng.sample(f(x_T), f(kw["init_flow"]), tables, t_scaled, [a for a, _ in ab], [b for _, b in ab],
                       None if init_feat0 is None else f(init_feat0), out)
            feat = eng.feat_nhwc().permute(0, 3, 1, 2).clone()
        return out, feat

    def _graphed(self, eng, tables, t_scaled, a, b, docs, n_batch, dev):
        key = (id(eng), id(tables))
        if getattr(self, "_fast", None) is None or self._fast[0] != key:
            self._fast = (key, _GraphedDocument(eng, tables, t_scaled, a, b, docs, n_batch, dev))
        return self._fast[1]


class _GraphedDocument:
    """Static input buffers + one CUDA graph of dvd_static_forward + dvd_sample for a fixed (engine, tables) pair."""

    def __init__(self, eng, tables, t_scaled, a, b, docs, n_batch, dev):
        self.eng, self.tables, self.t_scaled, self.a, self.b = eng, tables, t_scaled, a, b
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self.buf = {"y512": z(docs, 3, 512, 512), "mask_cat": z(docs, 1, 512, 512), "mask_y512": z(docs, 384, 64, 64),
                    "line_msk": z(docs, 64, 64, 64), "x_T": z(docs * n_batch, 2, 64, 64), "init_flow": z(docs, 2, 64, 64)}
        self.out = z(docs, 2, 64, 64)
        self.graph = None

    def _enqueue(self):
        bf = self.buf
        self.eng.static_forward(bf["y512"], bf["mask_cat"], bf["mask_y512"], bf["line_msk"])
        self.eng.sample(bf["x_T"], bf["init_flow"], self.tables, self.t_scaled, self.a, self.b, None, self.out)

    def run(self, inputs) -> torch.Tensor:
        for k, dst in self.buf.items():
            src = inputs[k]
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{k}: expected shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src, non_blocking=True)
        if self.graph is None:
            self._enqueue()                                # eager warm-up (function attributes, driver entry points)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g
        self.graph.replay()
        return self.out.clone()


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
