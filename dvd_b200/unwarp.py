"""Unwarp operators with the reference's call surface.

* ``register_model2`` / ``SpatialTransformer2``  (datasets/utils/warping.py:14-73): generic bilinear
  ``grid_sample`` (align_corners=True, zeros padding) — ``reg_model_bilin([img, grid]) -> img``.
* ``dewarp_fullres(map64, photo)``: evaluation.py:300-306 + visualization_utils.py:75-77 fused into
  ONE bandwidth-bound kernel (no materialised full-resolution grid).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

AFFINE = 0.987          # evaluation.py:306


def _chk(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: dvd_b200 has no CPU path")


class SpatialTransformer2(nn.Module):
    def __init__(self, size=None, mode="bilinear"):
        super().__init__()
        if mode != "bilinear":
            raise NotImplementedError("only bilinear sampling is on the val_TDiff path")
        self.mode = mode

    @torch.no_grad()
    def forward(self, src, flow):
        _chk(src, "src"); _chk(flow, "flow")
        src = src.float().contiguous(); flow = flow.float().contiguous()
        B, Cc, H, W = src.shape
        Bo, two, Ho, Wo = flow.shape
        assert two == 2 and Bo == B
        out = torch.empty((B, Cc, Ho, Wo), dtype=torch.float32, device=src.device)
        with torch.cuda.device(src.device):
            _lib.check(_lib.lib().dvd_grid_sample_f32(_lib.ptr(src), _lib.ptr(flow), _lib.ptr(out), B, Cc, H, W, Ho, Wo,
                                                      _lib.stream_ptr()), "dvd_grid_sample_f32")
        return out


class register_model2(nn.Module):                    # noqa: N801 (reference name)
    def __init__(self, img_size=(64, 1024, 1024), mode="bilinear"):
        super().__init__()
        self.spatial_trans = SpatialTransformer2(img_size, mode)

    def forward(self, x):
        return self.spatial_trans(x[0], x[1])


@torch.no_grad()
def dewarp_fullres(map64: torch.Tensor, photo: torch.Tensor, out: torch.Tensor | None = None, out_uint8: bool = False,
                   affine: float = AFFINE) -> torch.Tensor:
    """map64 [B,2,h,w] fp32 displacement (sampler output); photo either fp32 NCHW [B,C,H,W] (values
    0..255, the reference's ``source_image_ori``) or uint8 NHWC [B,H,W,C].
    Returns fp32 NCHW, or uint8 NHWC (truncated like ``.astype(np.uint8)``) when ``out_uint8`` / uint8 input."""
    _chk(map64, "map64"); _chk(photo, "photo")
    map64 = map64.float().contiguous()
    l = _lib.lib()
    B, _, mh, mw = map64.shape
    with torch.cuda.device(photo.device):
        st = _lib.stream_ptr()
        if photo.dtype == torch.uint8:
            photo = photo.contiguous()
            Bp, H, W, Cc = photo.shape
            assert Bp == B
            out = torch.empty_like(photo) if out is None else out
            _lib.check(l.dvd_unwarp_u8(_lib.ptr(photo), _lib.ptr(map64), _lib.ptr(out), B, Cc, H, W, mh, mw, affine, st), "dvd_unwarp_u8")
            return out
        photo = photo.float().contiguous()
        Bp, Cc, H, W = photo.shape
        assert Bp == B
        if out_uint8:
            out = torch.empty((B, H, W, Cc), dtype=torch.uint8, device=photo.device) if out is None else out
            _lib.check(l.dvd_unwarp_f32_u8(_lib.ptr(photo), _lib.ptr(map64), _lib.ptr(out), B, Cc, H, W, mh, mw, affine, st),
                       "dvd_unwarp_f32_u8")
            return out
        out = torch.empty_like(photo) if out is None else out
        _lib.check(l.dvd_unwarp_f32(_lib.ptr(photo), _lib.ptr(map64), _lib.ptr(out), B, Cc, H, W, mh, mw, affine, st), "dvd_unwarp_f32")
        return out


@torch.no_grad()
def fullres_grid(map64: torch.Tensor, H: int, W: int, affine: float = AFFINE) -> torch.Tensor:
    """evaluation.py:301-306 materialised ([B,2,H,W]); used by parity tests and by callers that still want the grid."""
    _chk(map64, "map64")
    map64 = map64.float().contiguous()
    B, _, mh, mw = map64.shape
    grid = torch.empty((B, 2, H, W), dtype=torch.float32, device=map64.device)
    with torch.cuda.device(map64.device):
        _lib.check(_lib.lib().dvd_fullres_grid_f32(_lib.ptr(map64), _lib.ptr(grid), B, H, W, mh, mw, affine, _lib.stream_ptr()),
                   "dvd_fullres_grid_f32")
    return grid
