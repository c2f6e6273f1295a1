"""Stage-by-stage comparison of the CUDA path with the CPU oracle on the GPU box (diagnostic, prints
a table; the asserting version lives in tests/test_gpu_parity.py).  Usage:
    python tools/gpu_stage_check.py [fp32|bf16] [doc_id]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dvd_oracle as O, synth          # noqa: E402
import dvd_b200                                     # noqa: E402
from dvd_b200.model import DiT                      # noqa: E402


def err(name, got, ref):
    got = got.detach().float().cpu(); ref = ref.detach().float().cpu()
    d = (got - ref).abs()
    print(f"{name:14s} max|d| {d.max():.3e}  mean|d| {d.mean():.3e}  ref rms {ref.pow(2).mean().sqrt():.3e}  "
          f"rel {d.max() / (ref.abs().max() + 1e-12):.2e}", flush=True)


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    doc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(1234, live_only=True)
    inp = synth.make_doc_inputs(doc, H=96, W=128)
    dev = torch.device("cuda:0")
    model = DiT(precision=prec)
    model.load_state_dict(sd, strict=False)
    model.to(dev)
    eng = model.engine(1, 2)
    g = lambda k: inp[k].to(dev).contiguous()
    t0 = time.time()
    eng.static_forward(g("y512"), g("mask_cat"), g("mask_y512"), g("line_msk"))
    torch.cuda.synchronize()
    print("static_forward ok", time.time() - t0, flush=True)
    st = O.Static(sd, inp["y512"], inp["mask_cat"], inp["mask_y512"], inp["line_msk"])
    err("feat", eng.feat_nhwc().permute(0, 3, 1, 2), st.feat)
    err("cond", eng.tensor("cond").view(1, 1024, 384), st.cond)
    err("msk6", eng.tensor("msk6").view(1, 1024, 384), st.msk6)
    err("msk_line", eng.tensor("msk_line").view(1, 1024, 384), st.msk_line)
    # ---- first step (t = 666.67 -> embeds 2, init_feat = feat)
    sch = O.Schedule(3)
    tab = eng.tables([2.0, 1.0, 0.0])
    temb = O.t_embed(sd, torch.tensor([2.0]))
    err("t_emb", tab[0, :384], temb[0])
    p = "blocks.11."
    ada = torch.nn.functional.linear(torch.nn.functional.silu(temb), sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
    err("adaLN", tab[0, 384:384 + 2304], ada[0])
    x = inp["x_T"]
    flow0 = torch.zeros(2, 2, 64, 64)
    pred = torch.empty(2, 2, 64, 64, device=dev); xprev = torch.empty_like(pred)
    a, b = sch.ddim_ab(2)
    eng.denoise_step(x.to(dev), flow0.to(dev), None, True, tab[0], a, b, pred, xprev)
    torch.cuda.synchronize()
    print("denoise_step ok", flush=True)
    xe = O.patch_embed(sd, "obs", x)
    err("xe", eng.tensor("xe").view(2, 1024, 384), xe)
    feat2 = st.feat.expand(2, -1, -1, -1)
    r = O.patch_embed(sd, "r", torch.cat([flow0, feat2], 1))
    err("r", eng.tensor("r").view(2, 1024, 384), r)
    rep = lambda v: v.expand(2, -1, -1)
    temb2 = temb.expand(2, -1)
    x4, x3, x2, x1 = O.dit_block_para(sd, 11, xe, temb2, rep(st.cond), rep(st.msk6), rep(st.msk_line), r)
    if prec == "fp32":
        qn = torch.nn.functional.layer_norm(xe, (384,), None, None, 1e-6)
        W, B = sd[p + "cross_attn.in_proj_weight"], sd[p + "cross_attn.in_proj_bias"]
        err("q", eng.tensor("q").view(2, 1024, 384), torch.nn.functional.linear(qn, W[:384], B[:384]))
    xc = torch.cat([x1, x2, x3, x4], 2)
    # X currently holds the decoder output residual stream (after 6 layers); compare final pred instead
    dec = O.decoder(sd, xc.transpose(1, 2).contiguous().view(2, 1536, 32, 32))
    fin = O.unpatchify(O.final_layer2(sd, dec, temb2)) + flow0
    err("pred(step0)", pred, fin)
    err("x_prev", xprev, O.ddim_update(sch, 2, x, fin))
    # ---- full loop
    out_ref, rec, _ = O.sample(sd, inp, S=3, n_batch=2, record=True)
    diff = model  # noqa
    from dvd_b200.sampler import create_gaussian_diffusion
    dif = create_gaussian_diffusion(steps=3, noise_schedule="cosine", predict_xstart=True, rescale_timesteps=True,
                                    rescale_learned_sigmas=True, timestep_respacing="")
    kw = {"init_flow": inp["init_flow"], "src_feat": None, "src_64": None, "y512": inp["y512"], "tmode": "stage_1_dit_cross",
          "mask_cat": inp["mask_cat"], "init_feat": inp["init_feat"], "iter": True, "mask_y512": inp["mask_y512"],
          "line_msk": inp["line_msk"]}
    torch.cuda.synchronize(); t0 = time.time()
    out, _ = dif.ddim_sample_loop(model, (1, 2, 64, 64), clip_denoised=False, model_kwargs=kw, eta=0.0, n_batch=2,
                                  time_variant=True, x_T=inp["x_T"])
    torch.cuda.synchronize()
    print("sample wall", time.time() - t0)
    err("map(S=3)", out, out_ref)
    gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"sample_S3_doc{doc}.npz"))
    err("map vs golden", out, torch.from_numpy(gold["sample"]))
    d = (out.cpu() - torch.from_numpy(gold["sample"])).abs()
    print(f"px error at 2000 px: mean {float(d.mean()) * 999.5:.4f} max {float(d.max()) * 999.5:.4f}; at 4032: mean {float(d.mean()) * 2015.5:.4f} max {float(d.max()) * 2015.5:.4f}")
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        out, _ = dif.ddim_sample_loop(model, (1, 2, 64, 64), clip_denoised=False, model_kwargs=kw, eta=0.0, n_batch=2,
                                      time_variant=True, x_T=inp["x_T"])
    torch.cuda.synchronize()
    print("sample wall (warm, avg of 3)", (time.time() - t0) / 3)


if __name__ == "__main__":
    main()
