"""Deterministic synthetic weights and inputs for the DvD hot path: the shared workload definition.

Neither product code nor oracle: a self-contained generator (numpy / torch only, no reference import, travels to the GPU box) that
``bench.py`` (both arms), ``tests/``, ``tools/`` and ``__graft_entry__.smoke()`` use to build identical inputs and random-init weights.
``oracle/synth.py`` re-exports it for the oracle's own scripts.

Why it exists: the reference ships no weights (README.md:47-57 points at Google Drive) and its own
``initialize_weights()`` zero-initialises every adaLN layer and ``final_layer2.linear``
(train_settings/dvd/improved_diffusion/cross_model.py:535-545), which makes the model output
identically ``init_flow``.  Parity would be vacuous.  We therefore define the "random-init
weights" ourselves, key by key, from per-key seeded CPU generators, and load the SAME state dict
into the reference (``oracle/make_golden.py``), the CPU oracle and the CUDA path.

Key names / shapes follow the instantiated reference model
``DiT_models2['DiT-S/2'](input_size=64, in_channels=2, tv=True)``
(cross_model.py:361-459, cross_attn.py:399-458); ``make_golden.py`` asserts the two agree.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

HID = 384           # DiT-S hidden size (cross_model.py:766-767)
DEPTH = 12
DEC_D = 1536        # decoder d_model = 4 * 384 (cross_model.py:446)
DEC_INNER = 2048
DEC_LAYERS = 6
TOKENS = 1024


# ----------------------------------------------------------------------------- fixed tables
def sincos_pos_embed_2d(embed_dim: int = HID, grid_size: int = 32) -> torch.Tensor:
    """Fixed 2-D sin-cos table, restating cross_model.py:677-722 (w-grid first, sin‖cos halves)."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_size, grid_size)

    def one_d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([one_d(embed_dim // 2, grid[0]), one_d(embed_dim // 2, grid[1])], axis=1)
    return torch.from_numpy(emb).float().unsqueeze(0)            # [1, 1024, 384]


def satrn_sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """Interleaved sin/cos table of cross_attn.py:122-134 -> [n_position, d_hid]."""
    denom = torch.Tensor([1.0 / np.power(10000, 2 * (j // 2) / d_hid) for j in range(d_hid)]).view(1, -1)
    pos = torch.arange(n_position).unsqueeze(-1).float()
    tab = pos * denom
    tab[:, 0::2] = torch.sin(tab[:, 0::2])
    tab[:, 1::2] = torch.cos(tab[:, 1::2])
    return tab


# ----------------------------------------------------------------------------- state-dict spec
def state_dict_spec() -> "OrderedDict[str, tuple]":
    """key -> shape for every entry of the reference state dict (369 entries)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["noised_obs_pos_embed"] = (1, TOKENS, HID)
    for name, (co, ci) in {
        "pyramid.level_0.0": (64, 4), "pyramid.level_1.0": (64, 64),
        "pyramid.level_2.0": (128, 64), "pyramid.level_2.2": (128, 128),
        "pyramid.level_3.0": (256, 128), "pyramid.level_3.2": (256, 256), "pyramid.level_3.4": (256, 256),
    }.items():
        s[name + ".weight"] = (co, ci, 3, 3)
        s[name + ".bias"] = (co,)
    for name, c in (("obs", 2), ("r", 258), ("c", 256), ("m", 384), ("l", 64)):
        s[f"{name}_embedder.proj.weight"] = (HID, c, 2, 2)
        s[f"{name}_embedder.proj.bias"] = (HID,)
    s["t_embedder.mlp.0.weight"] = (HID, 256)
    s["t_embedder.mlp.0.bias"] = (HID,)
    s["t_embedder.mlp.2.weight"] = (HID, HID)
    s["t_embedder.mlp.2.bias"] = (HID,)
    for i in range(DEPTH):
        p = f"blocks.{i}."
        s[p + "attn.qkv.weight"] = (3 * HID, HID)
        s[p + "attn.qkv.bias"] = (3 * HID,)
        s[p + "attn.proj.weight"] = (HID, HID)
        s[p + "attn.proj.bias"] = (HID,)
        s[p + "mlp.fc1.weight"] = (4 * HID, HID)
        s[p + "mlp.fc1.bias"] = (4 * HID,)
        s[p + "mlp.fc2.weight"] = (HID, 4 * HID)
        s[p + "mlp.fc2.bias"] = (HID,)
        s[p + "adaLN_modulation.1.weight"] = (6 * HID, HID)
        s[p + "adaLN_modulation.1.bias"] = (6 * HID,)
        s[p + "cross_attn.in_proj_weight"] = (3 * HID, HID)
        s[p + "cross_attn.in_proj_bias"] = (3 * HID,)
        s[p + "cross_attn.out_proj.weight"] = (HID, HID)
        s[p + "cross_attn.out_proj.bias"] = (HID,)
    s["decoder.position_dec.h_position_encoder"] = (1, DEC_D, 32, 1)
    s["decoder.position_dec.w_position_encoder"] = (1, DEC_D, 1, 32)
    for hw in ("h", "w"):
        for j in (0, 2):
            s[f"decoder.position_dec.{hw}_scale.{j}.weight"] = (DEC_D, DEC_D, 1, 1)
            s[f"decoder.position_dec.{hw}_scale.{j}.bias"] = (DEC_D,)
    for i in range(DEC_LAYERS):
        p = f"decoder.layer_stack.{i}."
        s[p + "norm1.weight"] = (DEC_D,)
        s[p + "norm1.bias"] = (DEC_D,)
        for n in ("linear_q", "linear_k", "linear_v", "fc"):
            s[p + f"attn.{n}.weight"] = (DEC_D, DEC_D)
        s[p + "norm2.weight"] = (DEC_D,)
        s[p + "norm2.bias"] = (DEC_D,)
        for n, wshape, c in (("conv1", (DEC_INNER, DEC_D, 1, 1), DEC_INNER),
                             ("depthwise_conv", (DEC_INNER, 1, 3, 3), DEC_INNER),
                             ("conv2", (DEC_D, DEC_INNER, 1, 1), DEC_D)):
            q = p + f"feed_forward.{n}."
            s[q + "conv.weight"] = wshape
            s[q + "bn.weight"] = (c,)
            s[q + "bn.bias"] = (c,)
            s[q + "bn.running_mean"] = (c,)
            s[q + "bn.running_var"] = (c,)
            s[q + "bn.num_batches_tracked"] = ()
    s["decoder.layer_norm.weight"] = (DEC_D,)
    s["decoder.layer_norm.bias"] = (DEC_D,)
    s["final_layer2.linear.weight"] = (8, DEC_D)
    s["final_layer2.linear.bias"] = (8,)
    s["final_layer2.adaLN_modulation.1.weight"] = (2 * DEC_D, DEC_D)
    s["final_layer2.adaLN_modulation.1.bias"] = (2 * DEC_D,)
    return s


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31 - 1))
    return g


def make_state_dict(seed: int = 1234, live_only: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Non-degenerate 'random-init' weights, one seeded generator per key.

    ``live_only`` skips DiT blocks 0..10, which the reference executes but whose outputs it
    discards (cross_model.py:614-616 never re-assigns ``x``); the CUDA path and the hoisted oracle
    accept a state dict without them (the reference loads with strict=False, val_TDiff.py:79).
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in state_dict_spec().items():
        if live_only and key.startswith("blocks.") and not key.startswith(f"blocks.{DEPTH - 1}."):
            continue
        g = _gen(key, seed)

        def nrm(std, mean=0.0):
            return torch.randn(shape, generator=g) * std + mean

        if key == "noised_obs_pos_embed":
            v = sincos_pos_embed_2d()
        elif key.endswith("h_position_encoder"):
            v = satrn_sinusoid_table(32, DEC_D).transpose(0, 1).reshape(1, DEC_D, 32, 1).contiguous()
        elif key.endswith("w_position_encoder"):
            v = satrn_sinusoid_table(32, DEC_D).transpose(0, 1).reshape(1, DEC_D, 1, 32).contiguous()
        elif key.endswith("num_batches_tracked"):
            v = torch.zeros((), dtype=torch.int64)
        elif key.endswith("running_mean"):
            v = nrm(0.1)
        elif key.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif ".bn." in key or ".norm1." in key or ".norm2." in key or key.startswith("decoder.layer_norm"):
            v = nrm(0.1, 1.0) if key.endswith("weight") else nrm(0.05)
        elif key == "final_layer2.linear.weight":
            v = nrm(0.002)
        elif key == "final_layer2.linear.bias":
            v = nrm(0.01)
        elif "adaLN_modulation" in key:
            v = nrm(0.02)
        elif key.endswith("bias") or key.endswith("in_proj_bias"):
            v = nrm(0.02)
        elif key.startswith("pyramid."):
            v = nrm(math.sqrt(2.0 / (shape[0] * 9)))
        elif "depthwise_conv.conv" in key:
            v = nrm(0.3)
        else:  # Linear / 1x1 conv / patch-embed conv weights: Xavier-like
            fan_out = shape[0]
            fan_in = int(np.prod(shape[1:]))
            v = nrm(math.sqrt(2.0 / (fan_in + fan_out)))
        sd[key] = v.contiguous()
    return sd


# ----------------------------------------------------------------------------- synthetic inputs
def make_photo(H: int, W: int, seed: int, kind: str = "page") -> torch.Tensor:
    """Synthetic warped-document photo, fp32 [1,3,H,W], integer-valued 0..255 (the reference
    dataset yields ``source_image_ori`` as 0..255 floats, datasets/doc_dataset/doc_benchmark.py:77-90)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "noise":
        return torch.floor(torch.rand((1, 3, H, W), generator=g) * 256.0).clamp_(0, 255)
    ys = torch.linspace(0, 1, H).view(H, 1)
    xs = torch.linspace(0, 1, W).view(1, W)
    ph = torch.rand(8, generator=g) * 6.283
    shade = 0.75 + 0.2 * torch.sin(2.1 * xs + ph[0]) * torch.cos(1.7 * ys + ph[1])
    lines = 0.5 + 0.5 * torch.sin(ys * (H / 9.0) + 0.8 * torch.sin(3.0 * xs + ph[2]))
    text = 0.5 + 0.5 * torch.sin(xs * (W / 5.0) + ph[3] + 5.0 * ys)
    ink = 1.0 - 0.85 * (lines > 0.55).float() * (text > 0.35).float()
    chans = []
    for c in range(3):
        tint = 0.92 + 0.08 * math.sin(1.3 * c + float(ph[4]))
        chans.append(torch.floor((shade * ink * tint).clamp(0, 1) * 255.0))
    return torch.stack(chans, 0).unsqueeze(0).contiguous()


def make_doc_inputs(doc_id: int, H: int = 1500, W: int = 2000, photo_kind: str = "page", with_photo: bool = True):
    """All per-document tensors of the hot path (evaluation.py:106-115 model_kwargs + x_T + photo).

    x_T is the SECOND randn after seeding (gaussian_diffusion.py:559-569: the first [1,2,64,64]
    draw is discarded)."""
    seed = 1000 + doc_id
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    if with_photo:
        photo = make_photo(H, W, seed, photo_kind)
        out["photo"] = photo
        out["y512"] = (F.interpolate(photo, size=(512, 512), mode="bilinear", align_corners=False) / 255.0).contiguous()
    else:
        out["y512"] = (make_photo(512, 512, seed, photo_kind) / 255.0).contiguous()
    yy = torch.linspace(-1, 1, 512).view(512, 1)
    xx = torch.linspace(-1, 1, 512).view(1, 512)
    r = torch.sqrt((xx / 0.85) ** 2 + (yy / 0.8) ** 2)
    out["mask_cat"] = torch.sigmoid((1.0 - r) * 12.0).view(1, 1, 512, 512).contiguous()
    out["mask_y512"] = torch.randn((1, 384, 64, 64), generator=g)
    out["line_msk"] = torch.randn((1, 64, 64, 64), generator=g)
    out["init_flow"] = torch.zeros((1, 2, 64, 64))
    out["init_feat"] = torch.zeros((1, 256, 64, 64))
    gx = torch.Generator(device="cpu").manual_seed(2000 + doc_id)
    _ = torch.randn((1, 2, 64, 64), generator=gx)
    out["x_T"] = torch.randn((2, 2, 64, 64), generator=gx)
    return out


def make_map64(doc_id: int, kind: str = "smooth", amp: float = 0.05) -> torch.Tensor:
    """A 64x64 backward-map DISPLACEMENT field [1,2,64,64] (what the sampler returns).
    'smooth' = bicubic-upsampled 8x8 N(0,amp^2) (realistic near-identity); 'adversarial' = iid noise."""
    g = torch.Generator(device="cpu").manual_seed(3000 + doc_id)
    if kind == "smooth":
        coarse = torch.randn((1, 2, 8, 8), generator=g) * amp
        return F.interpolate(coarse, size=(64, 64), mode="bicubic", align_corners=True).clamp(-1, 1).contiguous()
    if kind == "adversarial":
        return (torch.rand((1, 2, 64, 64), generator=g) * 2.4 - 1.2).clamp(-1, 1).contiguous()
    if kind == "zero":
        return torch.zeros((1, 2, 64, 64))
    raise ValueError(kind)
