"""Drop-in replacement for the reference denoiser ``DiT`` (DiT-S/2, tv=True).

Mirrors the call surface of train_settings/dvd/improved_diffusion/cross_model.py:361-651 that
``val_TDiff.py`` and the sampler use: ``model(x, t, **model_kwargs) -> (x0, feat)``,
``load_state_dict(sd, strict=False)``, ``state_dict()``, ``to()``, ``cpu()``, ``eval()``,
``parameters()``.  All arithmetic runs in libdvd_b200 (hand-written sm_100a kernels); there is no
PyTorch/CPU fallback — calling the model without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict, namedtuple

import torch
import torch.nn as nn

from . import _lib
from .weights import PackedWeights, required_keys

_IncompatibleKeys = namedtuple("IncompatibleKeys", ["missing_keys", "unexpected_keys"])

# fp32   : FFMA reference mode (bit-reproducible, no tensor cores)
# bf16x3 : tensor cores, fp32-accurate (split-bf16 operands, three tcgen05 passes per GEMM; fp16 attention) - the default: it is the
#          fastest mode that meets every accuracy gate of the reference comparison (map <= 0.05 px mean, image PSNR >= 45 dB)
# bf16   : tensor cores, single pass - about 2x faster again, map error ~0.6 px at 2000 px: NOT within the reference gates
PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "bf16x3": _lib.PREC_BF16X3}


def default_precision() -> str:
    return os.environ.get("DVD_PRECISION", "bf16x3")


def remap_t(t_scaled: float) -> float:
    """cross_model.py:575-579 (mode=None): strict thresholds, otherwise the raw value is embedded."""
    if t_scaled > 600:
        return 2.0
    if 600 > t_scaled > 300:
        return 1.0
    return float(t_scaled)


class Engine:
    """Device state for one (docs, n_hyp, precision) configuration: workspace + launch helpers."""

    def __init__(self, packed: PackedWeights, docs: int, n_hyp: int, precision: str):
        self.packed, self.docs, self.n_hyp = packed, docs, n_hyp
        self.prec = PRECISIONS[precision]
        self.lib = _lib.lib()
        _lib.check(self.lib.dvd_check_device(), "dvd_check_device")
        self.ws_bytes = int(self.lib.dvd_workspace_bytes(docs, n_hyp, self.prec))
        self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=packed.device)
        off = (-self.ws.data_ptr()) % 256
        self.ws_ptr = C.c_void_p(self.ws.data_ptr() + off)
        self.N = docs * n_hyp

    def tables(self, t_values) -> torch.Tensor:
        """dvd_tables_init for a list of (already remapped) timesteps -> [len, TABLE_ROW] device tensor.  Cached on the packed
        weights (the tables are a function of the weights and live on their device): a new state dict or .to() drops them."""
        key = tuple(float(v) for v in t_values)
        cache = self.packed.tables_cache
        if key not in cache:
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[key] = self._tables(t_values)
        return cache[key]

    def _tables(self, t_values) -> torch.Tensor:
        n = len(t_values)
        arr = (C.c_float * n)(*[float(v) for v in t_values])
        out = torch.empty((n, _lib.TABLE_ROW), dtype=torch.float32, device=self.packed.device)
        _lib.check(self.lib.dvd_tables_init(self.packed.ref(), arr, n, _lib.ptr(out), _lib.stream_ptr()), "dvd_tables_init")
        torch.cuda.current_stream().synchronize()      # `arr` is a host temporary
        return out

    def static_forward(self, y512, mask_cat, mask_y512, line_msk):
        for name, t, shp in (("y512", y512, (self.docs, 3, 512, 512)), ("mask_cat", mask_cat, (self.docs, 1, 512, 512)),
                             ("mask_y512", mask_y512, (self.docs, 384, 64, 64)), ("line_msk", line_msk, (self.docs, 64, 64, 64))):
            if tuple(t.shape) != shp or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise ValueError(f"{name}: expected contiguous CUDA fp32 {shp}, got {tuple(t.shape)} {t.dtype} {t.device}")
        _lib.check(self.lib.dvd_static_forward(self.packed.ref(), self.ws_ptr, self.ws_bytes, self.docs, self.n_hyp, self.prec,
                                               _lib.ptr(y512), _lib.ptr(mask_cat), _lib.ptr(mask_y512), _lib.ptr(line_msk),
                                               _lib.stream_ptr()), "dvd_static_forward")

    def denoise_step(self, x_t, init_flow, init_feat, feat_is_init, table_row, a, b, pred, x_prev):
        _lib.check(self.lib.dvd_denoise_step(self.packed.ref(), self.ws_ptr, self.ws_bytes, self.docs, self.n_hyp, self.prec,
                                             _lib.ptr(x_t), _lib.ptr(init_flow), _lib.ptr(init_feat), int(feat_is_init),
                                             _lib.ptr(table_row), float(a), float(b), _lib.ptr(pred), _lib.ptr(x_prev),
                                             _lib.stream_ptr()), "dvd_denoise_step")

    def sample(self, x_T, init_flow0, tables, t_scaled, ddim_a, ddim_b, init_feat0, map_out):
        S = len(t_scaled)
        fa = lambda v: (C.c_float * S)(*[float(x) for x in v])
        _lib.check(self.lib.dvd_sample(self.packed.ref(), self.ws_ptr, self.ws_bytes, self.docs, self.n_hyp, self.prec,
                                       _lib.ptr(x_T), _lib.ptr(init_flow0), _lib.ptr(tables), fa(t_scaled), fa(ddim_a), fa(ddim_b),
                                       S, _lib.ptr(init_feat0), _lib.ptr(map_out), _lib.stream_ptr()), "dvd_sample")

    def tensor(self, name: str) -> torch.Tensor:
        """Flat fp32 view of a named workspace buffer (stage-level parity tests)."""
        n = C.c_longlong(0)
        p = self.lib.dvd_workspace_tensor(self.ws_ptr, self.docs, self.n_hyp, self.prec, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        off = p - self.ws.data_ptr()
        return self.ws[off:off + 4 * n.value].view(torch.float32)

    def feat_nhwc(self) -> torch.Tensor:
        """View of the pyramid feature [docs,64,64,256] inside the workspace."""
        p = self.lib.dvd_workspace_feat(self.ws_ptr, self.docs, self.n_hyp, self.prec)
        off = p - self.ws.data_ptr()
        n = self.docs * 64 * 64 * 256
        return self.ws[off:off + 4 * n].view(torch.float32).view(self.docs, 64, 64, 256)


class DiT(nn.Module):
    """B200-native DiT-S/2 denoiser with the reference's interface (cross_model.py:361-651)."""

    def __init__(self, input_size=64, patch_size=2, in_channels=2, hidden_size=384, depth=12, num_heads=6, tv=True,
                 precision: str | None = None, **_ignored):
        super().__init__()
        if (input_size, patch_size, in_channels, hidden_size, depth, num_heads, bool(tv)) != (64, 2, 2, 384, 12, 6, True):
            raise NotImplementedError("dvd_b200 implements the val_TDiff configuration only: DiT-S/2, input 64, tv=True")
        self.in_channels, self.out_channels, self.patch_size, self.num_heads, self.tv = 2, 2, 2, 6, True
        self.precision = precision or default_precision()
        if self.precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {list(PRECISIONS)}")
        self._sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()      # master copy (CPU), reference key names
        self._device = torch.device("cpu")
        self._packed: PackedWeights | None = None
        self._engines: dict = {}
        self._anchor = nn.Parameter(torch.zeros(1), requires_grad=False)   # so that next(model.parameters()).device works

    # ------------------------------------------------------------------ nn.Module surface used by val_TDiff.py:79-85
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        need = required_keys()
        missing = [k for k in need if k not in state_dict]
        if strict and missing:
            raise RuntimeError(f"Missing key(s) in state_dict: {missing[:8]}{'...' if len(missing) > 8 else ''}")
        self._sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in state_dict.items())
        self._packed, self._engines = None, {}
        return _IncompatibleKeys(missing, [])

    def state_dict(self, *a, **k):
        return OrderedDict(self._sd)

    def parameters(self, recurse: bool = True):
        yield self._anchor
        for v in self._sd.values():
            if v.is_floating_point():
                yield v

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        dev = self._anchor.device
        if dev != self._device:
            self._device = dev
            self._packed, self._engines = None, {}
        return self

    @property
    def device(self):
        return self._anchor.device

    # ------------------------------------------------------------------ engine plumbing
    def packed(self) -> PackedWeights:
        if self._device.type != "cuda":
            raise RuntimeError("dvd_b200.DiT has no CPU path: move the model to a CUDA device (model.to('cuda'))")
        if not self._sd:
            raise RuntimeError("dvd_b200.DiT: load_state_dict() must be called before the first forward")
        if self._packed is None:
            with torch.cuda.device(self._device):
                self._packed = PackedWeights(self._sd, self._device, with_bf16=True)
        return self._packed

    def engine(self, docs: int, n_hyp: int, precision: str | None = None) -> Engine:
        precision = precision or self.precision
        key = (docs, n_hyp, precision)
        if key not in self._engines:
            self._engines = {k: v for k, v in self._engines.items() if k[2] != precision}   # one workspace per precision
            with torch.cuda.device(self._device):
                self._engines[key] = Engine(self.packed(), docs, n_hyp, precision)
        return self._engines[key]

    # ------------------------------------------------------------------ reference-compatible single forward
    @torch.no_grad()
    def forward(self, x, t, y=None, y512=None, mask_y512=None, init_flow=None, local_corr=None, trg_feat=None, src_feat=None,
                src_64=None, mask_x=None, tv=None, source_0=None, tmode=None, line_msk=None, mask_cat=None, init_feat=None,
                iter=False, mode=None):
        """cross_model.py:568-647.  Every tensor kwarg carries the sample batch N (the reference sampler
        repeats them, gaussian_diffusion.py:574); each sample is treated as its own document here.  The
        fused sampler (dvd_b200.sampler) hoists the per-document work instead of calling this."""
        if src_feat is not None or mask_y512 is None or line_msk is None or mask_cat is None or init_flow is None:
            raise NotImplementedError("dvd_b200.DiT supports the val_TDiff kwargs only (mask_y512, mask_cat, line_msk, init_flow)")
        if tv is not True:
            raise NotImplementedError("dvd_b200.DiT requires tv=True (time_variant)")
        if self._device.type != "cuda":
            raise RuntimeError("dvd_b200.DiT has no CPU path: move the model to a CUDA device (model.to('cuda'))")
        N = x.shape[0]
        f = lambda v: v.to(device=self._device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self._device):
            eng = self.engine(N, 1)
            eng.static_forward(f(y512), f(mask_cat), f(mask_y512), f(line_msk))
            t0 = float(t[0])
            tval = float(t0) if mode is not None else remap_t(t0)
            tab = eng.tables([tval])
            feat_is_init = bool(iter is True and (t0 > 600 or (N > 1 and bool((t == 2).all()))))     # cross_model.py:597-601
            if not feat_is_init and init_feat is None:
                raise ValueError("init_feat is required when t <= 600")
            pred = torch.empty((N, 2, 64, 64), dtype=torch.float32, device=self._device)
            eng.denoise_step(f(x), f(init_flow), None if feat_is_init else f(init_feat), feat_is_init, tab[0], 1.0, 0.0, pred, None)
            feat = eng.feat_nhwc().permute(0, 3, 1, 2).clone()              # logical NCHW, detached from the workspace
        return pred, feat


def DiT_S_2(**kwargs):
    return DiT(depth=12, hidden_size=384, patch_size=2, num_heads=6, **kwargs)


DiT_models2 = {"DiT-S/2": DiT_S_2}       # cross_model.py:779-784 (only the variant val_TDiff selects, script_util.py:155-162)
