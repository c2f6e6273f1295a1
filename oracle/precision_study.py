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
        return TF.linear(rnd(x, amode(m)), rnd(w, wmode(m)), b)

    @staticmethod
    def conv2d(x, w, b=None, **kw):
        m = Cfg.mode()
        if kw.get("groups", 1) > 1:        # depthwise conv is an elementwise kernel in fp32 in every mode
            return TF.conv2d(x, w, b, **kw)
        return TF.conv2d(rnd(x, amode(m)), rnd(w, wmode(m)), b, **kw)


def amode(m):                              # "a16w3": activations single fp16, weights as a bf16 pair (two passes); "ab16w3": single bf16
    return {"a16w3": "fp16", "ab16w3": "bf16"}.get(m, m)


def wmode(m):
    return {"a16w3": "x3", "ab16w3": "x3"}.get(m, m)


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


if __name__ == "__main__" and not any(f in sys.argv for f in ("--ln-fusion", "--qkv", "--two-pass", "--decoder-breakdown")):
    main()


# ----------------------------------------------------------------------------------------------- LayerNorm-fusion study
# Would feeding the RAW residual stream (as a bf16 pair) to the consumer GEMM and applying the normalisation in its epilogue,
#   LN(x) W^T = rstd * (x W'^T - mean * colsum(W')) + (beta W^T),   W' = W * gamma,
# keep the accuracy?  The operand error is then relative to |x|, not to |x - mean| / std.
def ln_linear_fused(x, w_ln, b_ln, eps, W, mode="x3", mod=None):
    """x [..., C] raw; returns LN(x)(*w_ln + b_ln)[*(1+scale)+shift] @ W^T computed the fused way with operands rounded by `mode`."""
    C = x.shape[-1]
    mean = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mean * mean                # single-pass statistics (sum, sum of squares)
    rstd = torch.rsqrt(var + eps)
    g = torch.ones(C) if w_ln is None else w_ln
    b = torch.zeros(C) if b_ln is None else b_ln
    if mod is not None:                                               # adaLN: (LN*g + b) * (1 + scale) + shift, per sample
        shift, scale = mod
        outs = []
        for n in range(x.shape[0]):
            gg, bb = g * (1 + scale[n]), b * (1 + scale[n]) + shift[n]
            Wp = W * gg[None, :]
            acc = TF.linear(rnd(x[n], mode), rnd(Wp, mode))
            outs.append(rstd[n] * (acc - mean[n] * Wp.sum(1)[None, :]) + (W @ bb)[None, :])
        return torch.stack(outs)
    Wp = W * g[None, :]
    acc = TF.linear(rnd(x, mode), rnd(Wp, mode))
    return rstd * (acc - mean * Wp.sum(1)) + (W @ b)


def decoder_fused(sd, x, mode="x3"):
    """oracle.decoder with LN1 -> q|k|v and LN2 -> conv1 computed the fused way (everything else as in the 'all x3' case)."""
    N, C, H, W = x.shape
    avg = x.mean(dim=(2, 3), keepdim=True)
    pd = "decoder.position_dec."

    def scale(hw):
        h = TF.relu(TF.conv2d(avg, sd[pd + f"{hw}_scale.0.weight"], sd[pd + f"{hw}_scale.0.bias"]))
        return torch.sigmoid(TF.conv2d(h, sd[pd + f"{hw}_scale.2.weight"], sd[pd + f"{hw}_scale.2.bias"]))
    x = x + scale("h") * sd[pd + "h_position_encoder"][:, :, :H, :] + scale("w") * sd[pd + "w_position_encoder"][:, :, :, :W]
    out = x.view(N, C, H * W).permute(0, 2, 1).contiguous()
    lin = lambda a, w: TF.linear(rnd(a, mode), rnd(w, mode))
    for i in range(6):
        p = f"decoder.layer_stack.{i}."
        Wqkv = torch.cat([sd[p + "attn.linear_q.weight"], sd[p + "attn.linear_k.weight"], sd[p + "attn.linear_v.weight"]], 0)
        qkv = ln_linear_fused(out, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5, Wqkv, mode)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        o = O._mha_core(q, k, v, O.HEADS, 1.0 / 16.0)
        out = out + lin(o, sd[p + "attn.fc.weight"])
        f = p + "feed_forward."
        h = ln_linear_fused(out, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5, sd[f + "conv1.conv.weight"].view(2048, C), mode)
        h = h.transpose(1, 2).contiguous().view(N, 2048, H, W)
        h = TF.relu(O._bn(sd, f + "conv1.", h))
        h = TF.relu(O._bn(sd, f + "depthwise_conv.", TF.conv2d(h, sd[f + "depthwise_conv.conv.weight"], padding=1, groups=2048)))
        h = TF.relu(O._bn(sd, f + "conv2.", TF.conv2d(rnd(h, mode), rnd(sd[f + "conv2.conv.weight"], mode))))
        out = out + h.view(N, C, H * W).transpose(1, 2)
    return O._ln(out, 1e-5, sd["decoder.layer_norm.weight"], sd["decoder.layer_norm.bias"])


def ln_fusion_study():
    torch.set_num_threads(8)
    sd = synth.make_state_dict(1234, live_only=True)
    inp = synth.make_doc_inputs(0, H=192, W=256)
    inp.pop("photo")
    patch()
    ref = run(sd, inp, {})
    groups = ["pyramid", "embed", "dit", "dit_attn", "decoder", "decoder_attn"]
    x3 = {**{g: "x3" for g in groups}, "dit_attn": "fp16", "decoder_attn": "fp16"}
    base = run(sd, inp, x3)
    e = (base - ref).abs()
    print(f"{'x3 + fp16 attention (shipping mode)':44s} mean {float(e.mean()):.3e} max {float(e.max()):.3e}  ({float(e.mean()) * 2015.5:.4f} px mean @4032)")
    dec = O.decoder
    O.decoder = lambda sd_, x: decoder_fused(sd_, x, "x3")
    try:
        out = run(sd, inp, x3)
    finally:
        O.decoder = dec
    e = (out - ref).abs()
    print(f"{'... with decoder LN1/LN2 fused into the GEMMs':44s} mean {float(e.mean()):.3e} max {float(e.max()):.3e}  ({float(e.mean()) * 2015.5:.4f} px mean @4032)")
    # how far from zero-mean are the decoder's rows?
    stats = []
    O.decoder = lambda sd_, x: (stats.append(x), dec(sd_, x))[1]
    try:
        run(sd, inp, {})
    finally:
        O.decoder = dec
    x = stats[0].flatten(2).transpose(1, 2)
    print(f"decoder input rows: |mean| / std  median {float((x.mean(-1).abs() / x.std(-1)).median()):.3f}  max {float((x.mean(-1).abs() / x.std(-1)).max()):.3f}")


if __name__ == "__main__" and "--ln-fusion" in sys.argv:
    ln_fusion_study()


# ----------------------------------------------------------------------------------------------- attention-input GEMM study
# q, k, v are rounded to fp16 before the attention MMAs anyway (2^-12), so do the GEMMs that PRODUCE them need all three split passes?
# Cases: the decoder's q|k|v GEMMs (18 of them, the largest GEMM of the step) with only one cross term kept (two passes: either the
# activation or the weight is a single bf16), or none (one pass).
def qkv_study():
    torch.set_num_threads(8)
    sd = synth.make_state_dict(1234, live_only=True)
    inp = synth.make_doc_inputs(0, H=192, W=256)
    inp.pop("photo")
    patch()
    ref = run(sd, inp, {})
    groups = ["pyramid", "embed", "dit", "dit_attn", "decoder", "decoder_attn"]
    x3 = {**{g: "x3" for g in groups}, "dit_attn": "fp16", "decoder_attn": "fp16"}
    qkv_ids = {id(sd[f"decoder.layer_stack.{i}.attn.linear_{n}.weight"]) for i in range(6) for n in "qkv"}
    special = {"x": "x3", "w": "x3"}

    def linear(x, w, b=None):
        if id(w) in qkv_ids:
            return TF.linear(rnd(x, special["x"]), rnd(w, special["w"]), b)
        m = Cfg.mode()
        return TF.linear(rnd(x, m), rnd(w, m), b)
    O.F.linear = linear
    print(f"{'decoder q|k|v GEMM operands':52s} {'mean':>10s} {'max':>10s} {'mean px@4032':>13s} {'max px@4032':>12s}")
    for name, xm, wm in (("x pair, w pair (3 passes, shipping)", "x3", "x3"), ("x pair, w bf16 (2 passes)", "x3", "bf16"),
                         ("x bf16, w pair (2 passes)", "bf16", "x3"), ("x bf16, w bf16 (1 pass)", "bf16", "bf16"),
                         ("x fp16, w fp16 (1 pass, fp16 operands)", "fp16", "fp16")):
        special["x"], special["w"] = xm, wm
        out = run(sd, inp, x3)
        e = (out - ref).abs()
        print(f"{name:52s} {float(e.mean()):10.3e} {float(e.max()):10.3e} {float(e.mean()) * 2015.5:13.4f} {float(e.max()) * 2015.5:12.4f}")
        sys.stdout.flush()


if __name__ == "__main__" and "--qkv" in sys.argv:
    qkv_study()


# ----------------------------------------------------------------------------------------------- two-pass study
# The weight rounding is what hurts (it is the same perturbation for every token and every step); the activation rounding is noise.
# Two passes = single 16-bit activation x (weight hi + weight lo).
def two_pass_study():
    torch.set_num_threads(8)
    sd = synth.make_state_dict(1234, live_only=True)
    inp = synth.make_doc_inputs(0, H=192, W=256)
    inp.pop("photo")
    patch()
    ref = run(sd, inp, {})
    gg = ["pyramid", "embed", "dit", "decoder"]
    attn = {"dit_attn": "fp16", "decoder_attn": "fp16"}
    cases = [("all x3 (3 passes, shipping)", {**{g: "x3" for g in gg}, **attn}),
             ("all: fp16 activation x weight pair (2 passes)", {**{g: "a16w3" for g in gg}, **attn}),
             ("all: bf16 activation x weight pair (2 passes)", {**{g: "ab16w3" for g in gg}, **attn})]
    cases += [(f"only {g}: fp16 activation x weight pair", {**{h: "x3" for h in gg}, g: "a16w3", **attn}) for g in gg]
    print(f"{'case':52s} {'mean':>10s} {'max':>10s} {'mean px@4032':>13s} {'max px@4032':>12s}")
    for name, modes in cases:
        out = run(sd, inp, modes)
        e = (out - ref).abs()
        print(f"{name:52s} {float(e.mean()):10.3e} {float(e.max()):10.3e} {float(e.mean()) * 2015.5:13.4f} {float(e.max()) * 2015.5:12.4f}")
        sys.stdout.flush()


if __name__ == "__main__" and "--two-pass" in sys.argv:
    two_pass_study()


def decoder_breakdown():
    """Which of the decoder's GEMMs tolerate a single fp16 activation operand (weights stay pairs)?"""
    torch.set_num_threads(8)
    sd = synth.make_state_dict(1234, live_only=True)
    inp = synth.make_doc_inputs(0, H=192, W=256)
    inp.pop("photo")
    patch()
    ref = run(sd, inp, {})
    gg = ["pyramid", "embed", "dit", "decoder"]
    x3 = {**{g: "x3" for g in gg}, "dit_attn": "fp16", "decoder_attn": "fp16"}
    ids = {"qkv": {id(sd[f"decoder.layer_stack.{i}.attn.linear_{n}.weight"]) for i in range(6) for n in "qkv"},
           "fc": {id(sd[f"decoder.layer_stack.{i}.attn.fc.weight"]) for i in range(6)},
           "conv1": {id(sd[f"decoder.layer_stack.{i}.feed_forward.conv1.conv.weight"]) for i in range(6)},
           "conv2": {id(sd[f"decoder.layer_stack.{i}.feed_forward.conv2.conv.weight"]) for i in range(6)}}
    chosen = set()

    def pick(x, w):
        if id(w) in chosen:
            return rnd(x, "fp16"), rnd(w, "x3")
        m = Cfg.mode()
        return rnd(x, amode(m)), rnd(w, wmode(m))

    def linear(x, w, b=None):
        a, ww = pick(x, w)
        return TF.linear(a, ww, b)

    def conv2d(x, w, b=None, **kw):
        if kw.get("groups", 1) > 1:
            return TF.conv2d(x, w, b, **kw)
        a, ww = pick(x, w)
        return TF.conv2d(a, ww, b, **kw)
    O.F.linear, O.F.conv2d = linear, conv2d
    print(f"{'decoder GEMMs with a single fp16 activation operand':52s} {'mean':>10s} {'max':>10s} {'mean px@4032':>13s} {'max px@4032':>12s}")
    for names in (["qkv"], ["fc"], ["conv1"], ["conv2"], ["qkv", "conv1"], ["qkv", "conv1", "fc"]):
        chosen.clear()
        for n in names:
            chosen.update(ids[n])
        out = run(sd, inp, x3)
        e = (out - ref).abs()
        print(f"{'+'.join(names):52s} {float(e.mean()):10.3e} {float(e.max()):10.3e} {float(e.mean()) * 2015.5:13.4f} {float(e.max()) * 2015.5:12.4f}")
        sys.stdout.flush()
    chosen.clear()
    chosen.update(ids["qkv"] | ids["conv1"])
    out = run(sd, inp, {**x3, "pyramid": "a16w3", "embed": "a16w3", "dit": "a16w3"})
    e = (out - ref).abs()
    print(f"{'qkv+conv1 + pyramid, embeds and DiT block':52s} {float(e.mean()):10.3e} {float(e.max()):10.3e} {float(e.mean()) * 2015.5:13.4f} {float(e.max()) * 2015.5:12.4f}")


if __name__ == "__main__" and "--decoder-breakdown" in sys.argv:
    decoder_breakdown()
