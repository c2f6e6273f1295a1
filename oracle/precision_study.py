"""Precision sensitivity study on the CPU oracle (TEST INFRASTRUCTURE, not product code).

Emulates reduced-precision tensor-core operands inside selected layer groups of the oracle (operands rounded to bf16, or to
a two-term bf16 split hi+lo = what the three-pass ``bf16x3`` tensor mode of libdvd_b200 multiplies exactly) while everything
else stays fp32, and reports the error of the final 64x64 backward map against the fp32 oracle run.  It answers which layers
may stay single-pass bf16 and what the split mode buys.  Output committed as profiles/r2_precision_study.txt.

    python -m oracle.precision_study [--doc 0]
"""
from __future__ import annotations

import argparse
import sys
import time

import torch
import torch.nn.functional as TF

from oracle import dvd_oracle as O
import synth_workload as synth


def rnd(x, mode):
    if mode == "fp32":
        return x
    if mode == "bf16":
        return x.bfloat16().float()
    if mode == "fp16":
        return x.half().float()
    if mode == "x3":                       # hi + lo, both bf16 (the lo*lo product is kept here; it is 2^-16 of the result)
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()
    raise ValueError(mode)


class Cfg:
    group = "other"
    modes = {}                             # group -> mode

    @classmethod
    def mode(cls):
        return cls.modes.get(cls.group, "fp32")


class FProxy:
    """Stands in for torch.nn.functional inside the oracle: rounds the operands of linear / conv2d."""

    def __getattr__(self, name):
        return getattr(TF, name)

    @staticmethod
    def linear(x, w, b=None):
        m = Cfg.mode()
        return TF.linear(rnd(x, m), rnd(w, m), b)

    @staticmethod
    def conv2d(x, w, b=None, **kw):
        m = Cfg.mode()
        if kw.get("groups", 1) > 1:        # depthwise conv is an elementwise kernel in fp32 in every mode
            return TF.conv2d(x, w, b, **kw)
        return TF.conv2d(rnd(x, m), rnd(w, m), b, **kw)


_orig = {}


def patch():
    O.F = FProxy()
    for name, group in (("pyramid", "pyramid"), ("patch_embed", "embed"), ("dit_block_para", "dit"), ("decoder", "decoder")):
        fn = getattr(O, name)
        _orig[name] = fn

        def wrap(*a, _fn=fn, _g=group, **k):
            prev = Cfg.group
            Cfg.group = _g
            try:
                return _fn(*a, **k)
            finally:
                Cfg.group = prev
        setattr(O, name, wrap)

    def mha_core(q, k, v, heads, scale):
        m = Cfg.modes.get(Cfg.group + "_attn", "fp32")
        N, Tq, C = q.shape
        d = C // heads
        qh = rnd(q, m).view(N, Tq, heads, d).transpose(1, 2)
        kh = rnd(k, m).view(N, -1, heads, d).transpose(1, 2)
        vh = rnd(v, m).view(N, -1, heads, d).transpose(1, 2)
        s = torch.matmul(qh, kh.transpose(2, 3)) * scale
        p = torch.exp(s - s.amax(-1, keepdim=True))
        pr = rnd(p, m)                      # the tensor core multiplies the rounded P; the row sum uses the same values
        return (torch.matmul(pr, vh) / pr.sum(-1, keepdim=True)).transpose(1, 2).reshape(N, Tq, C)
    O._mha_core = mha_core


def run(sd, inp, modes):
    Cfg.modes = modes
    with torch.no_grad():
        return O.sample(sd, inp, S=3, n_batch=2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--doc", type=int, default=0)
    a = ap.parse_args()
    torch.set_num_threads(8)
    sd = synth.make_state_dict(1234, live_only=True)
    inp = synth.make_doc_inputs(a.doc, H=192, W=256)
    inp.pop("photo")
    patch()
    ref = run(sd, inp, {})
    groups = ["pyramid", "embed", "dit", "dit_attn", "decoder", "decoder_attn"]
    cases = [("all bf16", {g: "bf16" for g in groups}), ("all fp16", {g: "fp16" for g in groups}), ("all x3", {g: "x3" for g in groups})]
    cases += [(f"only {g} bf16", {g: "bf16"}) for g in groups]
    cases += [("x3 but pyramid bf16", {**{g: "x3" for g in groups}, "pyramid": "bf16"}),
              ("x3 but pyramid+embed bf16", {**{g: "x3" for g in groups}, "pyramid": "bf16", "embed": "bf16"}),
              ("x3 but dit+dit_attn bf16", {**{g: "x3" for g in groups}, "dit": "bf16", "dit_attn": "bf16"}),
              ("x3 but attn (both) bf16", {**{g: "x3" for g in groups}, "dit_attn": "bf16", "decoder_attn": "bf16"}),
              ("x3 but attn (both) fp16", {**{g: "x3" for g in groups}, "dit_attn": "fp16", "decoder_attn": "fp16"})]
    print(f"doc {a.doc}: final-map error vs the fp32 oracle (normalised units; px = x (W-1)/2 with W = 2000 / 4032)")
    print(f"{'case':34s} {'mean':>10s} {'max':>10s} {'mean px@2000':>13s} {'max px@2000':>12s} {'mean px@4032':>13s} {'max px@4032':>12s}")
    for name, modes in cases:
        t0 = time.time()
        out = run(sd, inp, modes)
        e = (out - ref).abs()
        me, mx = float(e.mean()), float(e.max())
        print(f"{name:34s} {me:10.3e} {mx:10.3e} {me * 999.5:13.4f} {mx * 999.5:12.4f} {me * 2015.5:13.4f} {mx * 2015.5:12.4f}   ({time.time() - t0:.0f} s)")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
