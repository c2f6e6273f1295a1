"""Pins oracle/dvd_oracle.py (the CPU restatement) against fixtures produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dvd_oracle as O
from oracle import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_known_answers_schedule_and_tables(golden_dir):
    ka = _load(golden_dir, "known_answers.npz")
    s3, s10 = O.Schedule(3), O.Schedule(10)
    np.testing.assert_allclose(s3.betas, ka["betas3"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(s3.acp, ka["acp3"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(s3.acp_prev, ka["acp_prev3"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(s10.betas, ka["betas10"], rtol=0, atol=1e-15)
    # SURVEY.md §3.2 probe constants
    np.testing.assert_allclose(s3.betas, [0.2571163706, 0.6682546796, 0.999], atol=1e-9)
    a = [s3.ddim_ab(i)[0] for i in range(3)]
    b = [s3.ddim_ab(i)[1] for i in range(3)]
    np.testing.assert_allclose(a, [1, 0.5719249356, 0.4828061827], atol=1e-9)
    np.testing.assert_allclose(b, [0, 0.5841283698, 0.8681806204], atol=1e-9)
    te = O.timestep_embedding(torch.tensor([0.0, 1.0, 2.0, 333.33334, 600.0]), 256).numpy()
    np.testing.assert_array_equal(te, ka["tstep_emb"])
    np.testing.assert_array_equal(synth.sincos_pos_embed_2d()[0, 33].numpy(), ka["pos_embed_row33"])
    np.testing.assert_allclose(ka["pos_embed_row33"][:3], [0.8414710, 0.7885930, 0.7348220], atol=1e-6)


def test_t_remap_strict_thresholds():
    assert O.remap_t(666.6667) == 2.0 and O.remap_t(333.3333) == 1.0 and O.remap_t(0.0) == 0.0
    assert O.remap_t(600.0) == 600.0 and O.remap_t(300.0) == 300.0          # cross_model.py:576-579 strict
    assert O.remap_t(700.0) == 2.0 and O.remap_t(500.0) == 1.0 and O.remap_t(200.0) == 200.0


def test_ddim_collapse_matches_written_update():
    sch = O.Schedule(3)
    g = torch.Generator().manual_seed(0)
    x, p = torch.randn(2, 2, 64, 64, generator=g), torch.randn(2, 2, 64, 64, generator=g)
    for i in range(3):
        a, b = sch.ddim_ab(i)
        ref = O.ddim_update(sch, i, x, p)
        assert (ref - (a * p + b * x)).abs().max() < 2e-6
    assert torch.equal(O.ddim_update(sch, 0, x, p), p)


def test_stage_known_answers_first_forward(golden_dir, state_dict_live):
    """One denoiser forward (doc 0, first step) stage by stage vs the hooked reference."""
    st = _load(golden_dir, "stages_doc0_step0.npz")
    g3 = _load(golden_dir, "sample_S3_doc0.npz")
    sd = state_dict_live
    inp = synth.make_doc_inputs(0, H=96, W=128)
    static = O.Static(sd, inp["y512"], inp["mask_cat"], inp["mask_y512"], inp["line_msk"])
    np.testing.assert_allclose(static.feat[:, ::16, ::4, ::4].numpy(), g3["feat_sub"][:1], atol=2e-5, rtol=1e-5)
    pos = sd["noised_obs_pos_embed"]
    np.testing.assert_allclose((static.cond - pos)[:, ::8, ::8].numpy(), st["c_embed"][:1], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose((static.msk6 - pos)[:, ::8, ::8].numpy(), st["m_embed"][:1], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose((static.msk_line - pos)[:, ::8, ::8].numpy(), st["l_embed"][:1], atol=2e-5, rtol=1e-5)
    temb = O.t_embed(sd, torch.full((2,), 2.0))
    np.testing.assert_allclose(temb.numpy(), st["t_emb"], atol=1e-6)
    x = inp["x_T"]
    xe = O.patch_embed(sd, "obs", x)
    feat2 = static.feat.expand(2, -1, -1, -1)
    r = O.patch_embed(sd, "r", torch.cat([inp["init_flow"].repeat(2, 1, 1, 1), feat2], 1))
    np.testing.assert_allclose((r - pos)[:, ::8, ::8].numpy(), st["r_embed"], atol=3e-5, rtol=1e-5)
    rep = lambda v: v.expand(2, -1, -1)
    outs = O.dit_block_para(sd, 11, xe, temb, rep(static.cond), rep(static.msk6), rep(static.msk_line), r)
    np.testing.assert_allclose(np.stack([v[:, ::8, ::8].numpy() for v in outs]), st["block11"], atol=1e-4, rtol=1e-4)
    x4, x3, x2, x1 = outs
    xc = torch.cat([x1, x2, x3, x4], 2).transpose(1, 2).contiguous().view(2, 1536, 32, 32)
    dec = O.decoder(sd, xc)
    np.testing.assert_allclose(dec[:, ::8, ::16].numpy(), st["decoder"], atol=2e-4, rtol=1e-4)
    fin = O.final_layer2(sd, dec, temb)
    np.testing.assert_allclose(fin.numpy(), st["final"], atol=2e-5, rtol=1e-4)
    pred = O.unpatchify(fin)
    np.testing.assert_allclose(pred.numpy(), g3["pred"][0], atol=2e-5)


@pytest.mark.parametrize("doc", [0, 1])
def test_sampling_S3_matches_reference(golden_dir, state_dict_live, doc):
    g = _load(golden_dir, f"sample_S3_doc{doc}.npz")
    inp = synth.make_doc_inputs(doc, H=96, W=128)
    np.testing.assert_array_equal(inp["x_T"].numpy(), g["x"][0])
    out, rec, feat = O.sample(state_dict_live, inp, S=3, n_batch=2, record=True)
    for s in range(3):
        assert np.abs(rec["pred"][s].numpy() - g["pred"][s]).max() < 1e-4, s
        assert np.abs(rec["x"][s].numpy() - g["x"][s]).max() < 1e-4, s
    assert abs(rec["init_feat_cs"][1] - g["init_feat_sum"][1]) < 1e-3 * abs(g["init_feat_sum"][1]) + 1.0
    # 0.05 px mean / 0.5 px max at 4032 px  ==  2.48e-5 / 2.48e-4 normalised
    d = np.abs(out.numpy() - g["sample"])
    assert d.mean() < 2.0e-5 and d.max() < 2.0e-4, (d.mean(), d.max())


@pytest.mark.slow
def test_sampling_S10_thresholds(golden_dir, state_dict_live):
    g = _load(golden_dir, "sample_S10_doc2.npz")
    inp = synth.make_doc_inputs(2, H=96, W=128)
    out, rec, _ = O.sample(state_dict_live, inp, S=10, n_batch=2, record=True)
    for s in (0, 3, 4, 6, 7, 9):          # t = 900, 600 (raw), 500, 300 (raw), 200, 0
        assert np.abs(rec["pred"][s].numpy() - g["pred"][s]).max() < 5e-4, s
    assert np.abs(out.numpy() - g["sample"]).max() < 5e-4


def test_training_rollout_matches_reference(golden_dir, state_dict_live):
    """the no-grad roll-out of training_losses_time_variant (gaussian_diffusion.py:647-780), golden from the
    unmodified reference (oracle/make_golden_rollout.py)."""
    g = _load(golden_dir, "rollout_doc0_t0.npz")
    with torch.no_grad():
        out = O.rollout(state_dict_live, synth.make_doc_inputs(0, H=96, W=128), S=3, timestep=0)
    assert float((out - torch.from_numpy(g["pred"])).abs().max()) < 2e-5


@pytest.mark.parametrize("case", ["sampled_page", "smooth_noise", "adversarial_noise", "zero_page_1ch"])
def test_unwarp_matches_reference(golden_dir, case):
    u = _load(golden_dir, "unwarp.npz")
    g3 = _load(golden_dir, "sample_S3_doc0.npz")
    maps = {"sampled_page": (torch.from_numpy(g3["sample"]), synth.make_photo(96, 128, 11, "page")),
            "smooth_noise": (synth.make_map64(0, "smooth"), synth.make_photo(120, 90, 12, "noise")),
            "adversarial_noise": (synth.make_map64(1, "adversarial"), synth.make_photo(64, 200, 13, "noise")),
            "zero_page_1ch": (synth.make_map64(2, "zero"), synth.make_photo(77, 131, 14, "page")[:, :1].contiguous())}
    m, photo = maps[case]
    H, W = photo.shape[-2:]
    grid = O.fullres_grid(m, H, W)
    np.testing.assert_allclose(grid.numpy(), u[case + "_grid"], atol=1e-6)
    img = O.unwarp(m, photo)
    np.testing.assert_allclose(img.numpy(), u[case + "_img"], atol=1e-3)
    assert (O.to_uint8_hwc(img) != u[case + "_u8"]).mean() < 1e-3
