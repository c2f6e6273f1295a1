"""Generates tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python -m oracle.make_golden            # ~3 min on 8 cores

The fixtures pin ``oracle/dvd_oracle.py`` (tests/test_oracle_golden.py) and are compared directly
with the CUDA path (tests/test_gpu_parity.py).  Inputs and weights are NOT stored: they are
regenerated from ``oracle/synth.py`` seeds on both sides.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sub(t: torch.Tensor, *steps) -> np.ndarray:
    idx = tuple(slice(None, None, s) for s in steps)
    return t[idx].contiguous().numpy().astype(np.float32)


def main():
    torch.set_num_threads(os.cpu_count())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    os.makedirs(OUT, exist_ok=True)
    sd = synth.make_state_dict(1234)
    model = ref_harness.build_reference_model(sd)
    ref_sd = model.state_dict()
    spec = synth.state_dict_spec()
    assert list(ref_sd.keys()) == list(spec.keys()), "state-dict key order/name mismatch vs reference"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(spec[k]), (k, v.shape, spec[k])
    # the fixed tables we synthesise must equal what the reference's own constructor builds
    from train_settings.dvd.improved_diffusion.cross_model import DiT_models2, get_2d_sincos_pos_embed
    from train_settings.dvd.improved_diffusion.cross_attn import Adaptive2DPositionalEncoding
    pe = torch.from_numpy(get_2d_sincos_pos_embed(384, 32)).float().unsqueeze(0)
    assert torch.equal(pe, sd["noised_obs_pos_embed"])
    a2d = Adaptive2DPositionalEncoding(d_hid=1536, n_height=32, n_width=32)
    assert torch.equal(a2d.h_position_encoder, sd["decoder.position_dec.h_position_encoder"])
    assert torch.equal(a2d.w_position_encoder, sd["decoder.position_dec.w_position_encoder"])

    # ---------------------------------------------------------------- sampling, S=3 (default config), docs 0 and 1
    for doc in (0, 1):
        inp = synth.make_doc_inputs(doc, H=96, W=128)
        sample, rec = ref_harness.reference_sample(model, inp, S=3, n_batch=2, seed=2000 + doc)
        assert torch.equal(rec.calls[0]["x"], inp["x_T"]), "x_T (second randn after seeding) mismatch"
        g = {"sample": sample.numpy(), "t": np.stack([c["t"].numpy() for c in rec.calls]),
             "pred": np.stack([c["pred"].numpy() for c in rec.calls]), "x": np.stack([c["x"].numpy() for c in rec.calls]),
             "feat_sub": sub(rec.calls[0]["feat"], 1, 16, 4, 4),
             "init_feat_sub": np.stack([sub(c["init_feat"], 1, 16, 4, 4) for c in rec.calls]),
             "init_feat_sum": np.array([float(c["init_feat"].double().sum()) for c in rec.calls]),
             "feat_sum": np.array(float(rec.calls[0]["feat"].double().sum()))}
        np.savez_compressed(os.path.join(OUT, f"sample_S3_doc{doc}.npz"), **g)
        print(f"doc{doc} S=3: map std {sample.std():.4f} range [{sample.min():.3f},{sample.max():.3f}]"
              f" pred0 std {rec.calls[0]['pred'].std():.4f}")
        if doc == 0:
            map0 = sample.clone()
            # ------------------------------------------------------- stage-level known answers of ONE forward (first step)
            stages = {}
            hooks = []

            def keep(name, f):
                def hook(_m, _i, o):
                    stages[name] = f(o)
                return hook
            hooks.append(model.c_embedder.register_forward_hook(keep("c_embed", lambda o: sub(o, 1, 8, 8))))
            hooks.append(model.m_embedder.register_forward_hook(keep("m_embed", lambda o: sub(o, 1, 8, 8))))
            hooks.append(model.l_embedder.register_forward_hook(keep("l_embed", lambda o: sub(o, 1, 8, 8))))
            hooks.append(model.r_embedder.register_forward_hook(keep("r_embed", lambda o: sub(o, 1, 8, 8))))
            hooks.append(model.obs_embedder.register_forward_hook(keep("obs_embed", lambda o: sub(o, 1, 8, 8))))
            hooks.append(model.t_embedder.register_forward_hook(keep("t_emb", lambda o: o.numpy().copy())))
            hooks.append(model.blocks[11].register_forward_hook(
                keep("block11", lambda o: np.stack([sub(v, 1, 8, 8) for v in o]))))          # x4,x3,x2,x1
            hooks.append(model.decoder.position_dec.register_forward_hook(keep("posenc", lambda o: sub(o, 1, 16, 4, 4))))
            hooks.append(model.decoder.layer_stack[0].register_forward_hook(keep("dec_layer0", lambda o: sub(o, 1, 8, 16))))
            hooks.append(model.decoder.register_forward_hook(keep("decoder", lambda o: sub(o, 1, 8, 16))))
            hooks.append(model.final_layer2.register_forward_hook(keep("final", lambda o: o.numpy().copy())))
            c0 = rec.calls[0]
            with torch.no_grad():
                rep = lambda v: v.repeat(2, 1, 1, 1)
                out, _ = model(c0["x"], torch.tensor([2, 2]).float() * (1000.0 / 3), init_flow=rep(inp["init_flow"]),
                               init_feat=rep(inp["init_feat"]), y512=rep(inp["y512"]), mask_cat=rep(inp["mask_cat"]),
                               mask_y512=rep(inp["mask_y512"]), line_msk=rep(inp["line_msk"]), tmode="stage_1_dit_cross",
                               iter=True, tv=True)
            for h in hooks:
                h.remove()
            assert torch.equal(out, c0["pred"])
            np.savez_compressed(os.path.join(OUT, "stages_doc0_step0.npz"), **stages)

    # ---------------------------------------------------------------- S=10: exercises the strict t thresholds (600.0 / 300.0)
    inp = synth.make_doc_inputs(2, H=96, W=128)
    sample, rec = ref_harness.reference_sample(model, inp, S=10, n_batch=2, seed=2002)
    np.savez_compressed(os.path.join(OUT, "sample_S10_doc2.npz"), sample=sample.numpy(),
                        t=np.stack([c["t"].numpy() for c in rec.calls]), pred=np.stack([c["pred"].numpy() for c in rec.calls]),
                        init_feat_sum=np.array([float(c["init_feat"].double().sum()) for c in rec.calls]))
    print(f"doc2 S=10: map std {sample.std():.4f}")

    # ---------------------------------------------------------------- unwarp (evaluation.py:300-306 + grid_sample)
    u = {}
    cases = [("sampled_page", map0, synth.make_photo(96, 128, 11, "page")),
             ("smooth_noise", synth.make_map64(0, "smooth"), synth.make_photo(120, 90, 12, "noise")),
             ("adversarial_noise", synth.make_map64(1, "adversarial"), synth.make_photo(64, 200, 13, "noise")),
             ("zero_page_1ch", synth.make_map64(2, "zero"), synth.make_photo(77, 131, 14, "page")[:, :1].contiguous())]
    for name, m, photo in cases:
        grid, img = ref_harness.reference_unwarp(m, photo)
        u[name + "_grid"] = grid.numpy()
        u[name + "_img"] = img.numpy()
        u[name + "_u8"] = img[0].permute(1, 2, 0).numpy().astype(np.uint8)       # visualization_utils.py:76-77
    np.savez_compressed(os.path.join(OUT, "unwarp.npz"), **u)

    # ---------------------------------------------------------------- small known answers
    from train_settings.dvd.improved_diffusion.cross_model import TimestepEmbedder
    d3 = ref_harness.build_reference_diffusion(3)
    d10 = ref_harness.build_reference_diffusion(10)
    np.savez(os.path.join(OUT, "known_answers.npz"),
             betas3=d3.betas, acp3=d3.alphas_cumprod, acp_prev3=d3.alphas_cumprod_prev,
             betas10=d10.betas, acp10=d10.alphas_cumprod,
             tstep_emb=TimestepEmbedder.timestep_embedding(torch.tensor([0.0, 1.0, 2.0, 333.33334, 600.0]), 256).numpy(),
             pos_embed_row33=pe[0, 33].numpy())
    print("golden written to", OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})


if __name__ == "__main__":
    main()
