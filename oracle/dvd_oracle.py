"""CPU ORACLE for the DvD sampling + unwarp hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

A functional restatement, in plain fp32 PyTorch CPU ops, of the reference's algorithm for the
path named by BASELINE.json:north_star.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``dvd_b200`` never does (it raises if its CUDA library is missing).

Parity pinning: the reference has no tests / golden vectors, so this oracle is
pinned against outputs of the UNMODIFIED reference run in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).

Every function cites the reference lines it follows (paths relative to /root/reference):
  GD  = train_settings/dvd/improved_diffusion/gaussian_diffusion.py
  RS  = train_settings/dvd/improved_diffusion/respace.py
  CM  = train_settings/dvd/improved_diffusion/cross_model.py
  CA  = train_settings/dvd/improved_diffusion/cross_attn.py
  EV  = train_settings/dvd/evaluation.py
  WP  = datasets/utils/warping.py
Third-party arithmetic restated from published semantics (absent from /root/reference):
  timm Attention/Mlp/PatchEmbed (unpinned, requirements.txt:17), mmcv ConvModule (2.2.0, comment
  in requirements.txt:4), torch nn.MultiheadAttention / F.grid_sample / F.interpolate.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

HEADS = 6


# ----------------------------------------------------------------------------- schedule (GD)
def cosine_betas(S: int, max_beta: float = 0.999) -> np.ndarray:
    """GD:49-75 get_named_beta_schedule('cosine') / betas_for_alpha_bar."""
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - ab((i + 1) / S) / ab(i / S), max_beta) for i in range(S)], dtype=np.float64)


def linear_betas(S: int) -> np.ndarray:
    """GD:41-48."""
    scale = 1000 / S
    return np.linspace(scale * 0.0001, scale * 0.02, S, dtype=np.float64)


class Schedule:
    """float64 tables of GD:171-212 (timestep_respacing='' => SpacedDiffusion keeps every step, RS:63-86)."""

    def __init__(self, S: int, name: str = "cosine"):
        betas = cosine_betas(S) if name == "cosine" else linear_betas(S)
        self.S = S
        self.betas = betas
        self.acp = np.cumprod(1.0 - betas)
        self.acp_prev = np.append(1.0, self.acp[:-1])
        self.sqrt_recip_acp = np.sqrt(1.0 / self.acp)
        self.sqrt_recipm1_acp = np.sqrt(1.0 / self.acp - 1)

    def scaled_t(self, i: int) -> float:
        """RS:111-123: new_ts = timestep_map[t].float() * (1000 / S); fp32 arithmetic."""
        return float(np.float32(i) * np.float32(1000.0 / self.S))

    def ddim_ab(self, i: int):
        """eta=0 collapse of GD:445-491: x_{t-1} = a*pred + b*x_t (float64)."""
        sp = math.sqrt(1.0 - self.acp_prev[i])
        a = math.sqrt(self.acp_prev[i]) - sp / self.sqrt_recipm1_acp[i]
        b = sp * self.sqrt_recip_acp[i] / self.sqrt_recipm1_acp[i]
        return a, b


def ddim_update(sch: Schedule, i: int, x: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """GD:470-489 with eta=0 (sigma=0, the drawn noise is multiplied by 0), op order as written."""
    f = lambda v: torch.tensor(v, dtype=torch.float64).float().to(x.device)
    eps = (f(sch.sqrt_recip_acp[i]) * x - pred) / f(sch.sqrt_recipm1_acp[i])
    abp = f(sch.acp_prev[i])
    return pred * torch.sqrt(abp) + torch.sqrt(1 - abp - 0.0) * eps


# ----------------------------------------------------------------------------- denoiser pieces (CM, CA)
def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: int = 10000) -> torch.Tensor:
    """CM:111-134 (cos ‖ sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def remap_t(t_scaled: float) -> float:
    """CM:575-579 (mode=None): strict inequalities; otherwise the raw scaled value is embedded."""
    if t_scaled > 600:
        return 2.0
    if 600 > t_scaled > 300:
        return 1.0
    return t_scaled


def t_embed(sd, t: torch.Tensor) -> torch.Tensor:
    """CM:136-139."""
    h = F.linear(timestep_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    return F.linear(F.silu(h), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])


def pyramid(sd, y4: torch.Tensor) -> torch.Tensor:
    """CM:18-95 VGGPyramid(64): [N,4,512,512] -> level_3 [N,256,64,64]."""
    c = lambda x, n: F.relu(F.conv2d(x, sd[f"pyramid.{n}.weight"], sd[f"pyramid.{n}.bias"], padding=1))
    x = c(y4, "level_0.0")
    x = F.max_pool2d(c(x, "level_1.0"), 2)
    x = F.max_pool2d(c(c(x, "level_2.0"), "level_2.2"), 2)
    x = F.max_pool2d(c(c(c(x, "level_3.0"), "level_3.2"), "level_3.4"), 2)
    return x


def patch_embed(sd, name: str, x: torch.Tensor) -> torch.Tensor:
    """timm PatchEmbed (CM:396-411): conv k=2,s=2 -> flatten(2).transpose(1,2); + pos (CM:571,585,594,603,605)."""
    y = F.conv2d(x, sd[f"{name}_embedder.proj.weight"], sd[f"{name}_embedder.proj.bias"], stride=2)
    return y.flatten(2).transpose(1, 2) + sd["noised_obs_pos_embed"]


def _ln(x, eps, w=None, b=None):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _mha_core(q, k, v, heads, scale):
    """softmax(q k^T * scale) v with [N,T,heads*d] layout in and out."""
    N, Tq, C = q.shape
    d = C // heads
    qh = q.view(N, Tq, heads, d).transpose(1, 2)
    kh = k.view(N, -1, heads, d).transpose(1, 2)
    vh = v.view(N, -1, heads, d).transpose(1, 2)
    att = torch.softmax(torch.matmul(qh * scale, kh.transpose(2, 3)), dim=-1)
    return torch.matmul(att, vh).transpose(1, 2).reshape(N, Tq, C)


def cross_mha(sd, p: str, q_in, kv_in):
    """torch nn.MultiheadAttention(batch_first) as used at CM:203-205,237-265 (packed in_proj)."""
    W, B = sd[p + "cross_attn.in_proj_weight"], sd[p + "cross_attn.in_proj_bias"]
    D = q_in.shape[-1]
    q = F.linear(q_in, W[:D], B[:D])
    k = F.linear(kv_in, W[D:2 * D], B[D:2 * D])
    v = F.linear(kv_in, W[2 * D:], B[2 * D:])
    o = _mha_core(q, k, v, HEADS, (D // HEADS) ** -0.5)
    return F.linear(o, sd[p + "cross_attn.out_proj.weight"], sd[p + "cross_attn.out_proj.bias"])


def self_attn(sd, p: str, x):
    """timm Attention (qkv bias, scale hd^-0.5) as called at CM:268-292."""
    D = x.shape[-1]
    qkv = F.linear(x, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    o = _mha_core(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], HEADS, (D // HEADS) ** -0.5)
    return F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])


def mlp(sd, p: str, x):
    """timm Mlp with GELU(tanh) (CM:167-174)."""
    h = F.gelu(F.linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]), approximate="tanh")
    return F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def modulate(x, shift, scale):
    """CM:13-14."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def dit_block_para(sd, i: int, x, temb, cond, msk6, msk_line, r):
    """CM:208-293 DiTBlock.forward, separate_cross_attn='para', tv=True -> (x4, x3, x2, x1)."""
    p = f"blocks.{i}."
    ada = F.linear(F.silu(temb), sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = ada.chunk(6, dim=1)
    q = _ln(x, 1e-6)
    outs = []
    for ctx in (cond, msk6, msk_line, r):
        xs = x + cross_mha(sd, p, q, ctx)
        xs = xs + g_a.unsqueeze(1) * self_attn(sd, p, modulate(_ln(xs, 1e-6), sh_a, sc_a))
        xs = xs + g_m.unsqueeze(1) * mlp(sd, p, modulate(_ln(xs, 1e-6), sh_m, sc_m))
        outs.append(xs)
    x1, x2, x3, x4 = outs
    return x4, x3, x2, x1


def _bn(sd, q, x):
    return F.batch_norm(x, sd[q + "bn.running_mean"], sd[q + "bn.running_var"], sd[q + "bn.weight"], sd[q + "bn.bias"],
                        training=False, eps=1e-5)


def decoder(sd, x):
    """CA:399-458 Decoder.forward on [N,1536,32,32] -> [N,1024,1536] (mask all ones: CA:443-451)."""
    N, C, H, W = x.shape
    # Adaptive2DPositionalEncoding CA:143-157
    avg = x.mean(dim=(2, 3), keepdim=True)
    pd = "decoder.position_dec."

    def scale(hw):
        h = F.relu(F.conv2d(avg, sd[pd + f"{hw}_scale.0.weight"], sd[pd + f"{hw}_scale.0.bias"]))
        return torch.sigmoid(F.conv2d(h, sd[pd + f"{hw}_scale.2.weight"], sd[pd + f"{hw}_scale.2.bias"]))

    x = x + scale("h") * sd[pd + "h_position_encoder"][:, :, :H, :] + scale("w") * sd[pd + "w_position_encoder"][:, :, :, :W]
    out = x.view(N, C, H * W).permute(0, 2, 1).contiguous()
    for i in range(6):
        p = f"decoder.layer_stack.{i}."
        # DecoderLayer CA:377-396 ; MultiHeadAttention CA:197-221 (no biases, temperature d_k**0.5=16)
        h = _ln(out, 1e-5, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        q = F.linear(h, sd[p + "attn.linear_q.weight"])
        k = F.linear(h, sd[p + "attn.linear_k.weight"])
        v = F.linear(h, sd[p + "attn.linear_v.weight"])
        o = _mha_core(q, k, v, HEADS, 1.0 / 16.0)
        out = out + F.linear(o, sd[p + "attn.fc.weight"])
        h = _ln(out, 1e-5, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        h = h.transpose(1, 2).contiguous().view(N, C, H, W)
        # LocalityAwareFeedforward CA:52-57 (ConvModule = conv(no bias) + BN(eval) + ReLU, incl. after conv2)
        f = p + "feed_forward."
        h = F.relu(_bn(sd, f + "conv1.", F.conv2d(h, sd[f + "conv1.conv.weight"])))
        h = F.relu(_bn(sd, f + "depthwise_conv.", F.conv2d(h, sd[f + "depthwise_conv.conv.weight"], padding=1, groups=h.shape[1])))
        h = F.relu(_bn(sd, f + "conv2.", F.conv2d(h, sd[f + "conv2.conv.weight"])))
        out = out + h.view(N, C, H * W).transpose(1, 2)
    return _ln(out, 1e-5, sd["decoder.layer_norm.weight"], sd["decoder.layer_norm.bias"])


def final_layer2(sd, x, temb):
    """CM:329-336 (tv=True: t.repeat(1,4))."""
    ada = F.linear(F.silu(temb.repeat(1, 4)), sd["final_layer2.adaLN_modulation.1.weight"], sd["final_layer2.adaLN_modulation.1.bias"])
    shift, scale = ada.chunk(2, dim=1)
    return F.linear(modulate(_ln(x, 1e-6), shift, scale), sd["final_layer2.linear.weight"], sd["final_layer2.linear.bias"])


def unpatchify(x, c: int = 2, p: int = 2):
    """CM:553-566 'nhwpqc->nchpwq'."""
    N, T, _ = x.shape
    h = w = int(T ** 0.5)
    x = x.reshape(N, h, w, p, p, c)
    return torch.einsum("nhwpqc->nchpwq", x).reshape(N, c, h * p, w * p)


class Static:
    """Per-document, step-invariant tensors (the reference recomputes them every forward)."""

    def __init__(self, sd, y512, mask_cat, mask_y512, line_msk):
        self.feat = pyramid(sd, torch.cat([y512, mask_cat], dim=1))   # CM:586-589
        self.cond = patch_embed(sd, "c", self.feat)                    # CM:594
        self.msk6 = patch_embed(sd, "m", mask_y512)                    # CM:585
        self.msk_line = patch_embed(sd, "l", line_msk)                 # CM:605


def denoiser_forward(sd, x, t_scaled: float, init_flow, init_feat, static: Static = None, *, y512=None, mask_cat=None,
                     mask_y512=None, line_msk=None, as_written: bool = False, raw_t: bool = False):
    """CM:568-647 DiT.forward (mode=None, tv=True, iter=True, src_feat=None) -> (x0, feat).

    ``static`` may hold tensors for ONE document; they are broadcast over the N hypotheses
    (GD:574 repeats every kwarg n_batch times, so this is exact).  ``as_written`` also runs the 11
    dead DiT blocks (CM:614-616) so that CPU-baseline timings reflect the reference's own work."""
    N = x.shape[0]
    if static is None:
        static = Static(sd, y512, mask_cat, mask_y512, line_msk)
    rep = lambda v: v.expand(N, *v.shape[1:]) if v.shape[0] != N else v
    xe = patch_embed(sd, "obs", x)                                               # CM:571
    # CM:575-580: the strict-threshold remap only when mode is None; the training roll-out (mode='train') embeds the raw value
    temb = t_embed(sd, torch.full((N,), t_scaled if raw_t else remap_t(t_scaled), dtype=torch.float32, device=x.device))
    feat = rep(static.feat)
    if t_scaled > 600 or (N > 1 and t_scaled == 2.0):                            # CM:597-601 (entries whose float t equals the label 2)
        init_feat = feat
    r = patch_embed(sd, "r", torch.cat([init_flow, init_feat], dim=1))           # CM:602-603
    blocks = range(12) if as_written else (11,)
    for i in blocks:                                                             # CM:614-616 (x never updated)
        x4, x3, x2, x1 = dit_block_para(sd, i, xe, temb, rep(static.cond), rep(static.msk6), rep(static.msk_line), r)
    d = x3.shape[-1]
    xc = torch.cat([x1, x2, x3, x4], dim=2).transpose(1, 2).contiguous().view(N, 4 * d, 32, 32)   # CM:623
    y = decoder(sd, xc)                                                          # CM:624
    y = final_layer2(sd, y, temb)                                                # CM:625-627
    y = unpatchify(y)                                                            # CM:644
    return y + init_flow, feat                                                   # CM:645-647


# ----------------------------------------------------------------------------- sampler (GD:537-645)
def base_grid(n: int) -> torch.Tensor:
    """coords_grid_tensor((n,n))/(n-1): channel 0 = x/column ramp, 1 = y/row ramp (GD:23-28,219-223)."""
    r = torch.linspace(0, n - 1, n) / (n - 1)
    return torch.stack([r.view(1, n).expand(n, n), r.view(n, 1).expand(n, n)], 0).unsqueeze(0).float()


def grid_sample_ref(src, grid_nchw):
    grid_nchw = grid_nchw.to(src.device)
    """WP:50-73 SpatialTransformer2.forward."""
    return F.grid_sample(src, grid_nchw.permute(0, 2, 3, 1), align_corners=True, mode="bilinear", padding_mode="zeros")


def sample(sd, inp: dict, S: int = 3, n_batch: int = 2, schedule: str = "cosine", as_written: bool = False,
           record: bool = False):
    """ddim_sample_loop (GD:494-535) -> ddim_sample_loop_progressive_only_mean (GD:537-645), iter=True,
    time_variant=True, eta=0, followed by the clamp of EV:137.  ``inp`` as from ``synth.make_doc_inputs``."""
    sch = Schedule(S, schedule)
    img = inp["x_T"].clone()
    rep = lambda v: v.repeat(n_batch, 1, 1, 1)                                   # GD:574
    init_flow, init_feat = rep(inp["init_flow"]), rep(inp["init_feat"])
    static = None if as_written else Static(sd, inp["y512"], inp["mask_cat"], inp["mask_y512"], inp["line_msk"])
    b64 = base_grid(64).to(img.device)
    rec = {"pred": [], "x": [], "init_feat_cs": []}
    pred = feat = None
    for i in range(S - 1, -1, -1):                                               # GD:597
        if i != S - 1:                                                           # GD:618-624
            init_flow = pred.clone()
            init_feat = grid_sample_ref(feat, (init_flow + b64) * 2 - 1)
        if as_written:
            pred, feat = denoiser_forward(sd, img, sch.scaled_t(i), init_flow, init_feat, None, y512=rep(inp["y512"]),
                                          mask_cat=rep(inp["mask_cat"]), mask_y512=rep(inp["mask_y512"]),
                                          line_msk=rep(inp["line_msk"]), as_written=True)
        else:
            pred, feat = denoiser_forward(sd, img, sch.scaled_t(i), init_flow, init_feat, static)
        if record:
            rec["x"].append(img.clone()); rec["pred"].append(pred.clone()); rec["init_feat_cs"].append(float(init_feat.double().sum()))
        img = ddim_update(sch, i, img, pred)                                      # GD:470-489
    out = torch.clamp(pred.mean(dim=0, keepdim=True), -1, 1)                     # GD:639-640, EV:137
    return (out, rec, feat) if record else out


def rollout(sd, inp: dict, S: int = 3, timestep: int = 0, schedule: str = "cosine"):
    """ddim_sample_loop_for_training (GD:647-780) as called by training_losses_time_variant (GD:924-942): one sample, DDIM steps
    S-1 .. timestep+1, raw timestep embedding (mode='train'), returns clamp(pred_xstart).  ``inp['x_T'][:1]`` is the initial noise."""
    sch = Schedule(S, schedule)
    img = inp["x_T"][:1].clone()
    init_flow, init_feat = inp["init_flow"].clone(), inp["init_feat"].clone()
    static = Static(sd, inp["y512"], inp["mask_cat"], inp["mask_y512"], inp["line_msk"])
    b64 = base_grid(64)
    pred = feat = None
    for i in range(S - 1, timestep, -1):                                         # GD:723
        if i != S - 1:                                                           # GD:738-759
            init_flow = pred.clone()
            init_feat = grid_sample_ref(feat, (init_flow + b64) * 2 - 1)
        pred, feat = denoiser_forward(sd, img, sch.scaled_t(i), init_flow, init_feat, static, raw_t=True)
        img = ddim_update(sch, i, img, pred)
    return torch.clamp(pred, -1, 1)


# ----------------------------------------------------------------------------- unwarp (EV:300-306 + WP:73)
def fullres_grid(map64: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """EV:301-306: upsample map, upsample the 512^2 base ramp, affine 0.987."""
    s = F.interpolate(map64, size=(H, W), mode="bilinear", align_corners=True)
    base = F.interpolate(base_grid(512).to(map64.device), size=(H, W), mode="bilinear", align_corners=True)
    return ((s + base) * 1 * 2 - 1) * 0.987


def unwarp(map64: torch.Tensor, photo: torch.Tensor) -> torch.Tensor:
    """EV:300-318 -> visualization_utils.py:75 -> WP:73; returns fp32 [1,C,H,W] (pre-uint8)."""
    H, W = photo.shape[-2:]
    return grid_sample_ref(photo.float(), fullres_grid(map64, H, W))


def to_uint8_hwc(img: torch.Tensor) -> np.ndarray:
    """visualization_utils.py:76-77: permute -> numpy -> astype(uint8) (truncation)."""
    return img[0].permute(1, 2, 0).numpy().astype(np.uint8)
