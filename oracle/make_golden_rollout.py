"""Golden vector for the training-time roll-out: runs the UNMODIFIED reference's
``ddim_sample_loop_for_training`` (gaussian_diffusion.py:647-780) on the CPU for document 0, timestep = 0 (DDIM steps 2, 1) with
the x_T of the shared synthetic workload, and writes tests/golden/rollout_doc0_t0.npz.  TEST INFRASTRUCTURE (build container only).

    python -m oracle.make_golden_rollout
"""
import os

import numpy as np
import torch

from oracle import ref_harness as RH
import synth_workload as synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(1234)
    model = RH.build_reference_model(sd)
    inp = synth.make_doc_inputs(0, H=96, W=128)
    diffusion = RH.build_reference_diffusion(3)
    kwargs = {"init_flow": inp["init_flow"].clone(), "y512": inp["y512"], "mask_cat": inp["mask_cat"], "init_feat": inp["init_feat"].clone(),
              "mask_y512": inp["mask_y512"], "line_msk": inp["line_msk"]}
    # the reference draws randn(shape) then randn(n_batch, ...): feed the workload's first hypothesis noise through the same two draws
    x_T = inp["x_T"][:1].clone()
    orig = torch.randn
    draws = []

    def fake_randn(*a, **k):
        draws.append(a)
        return x_T.clone() if len(draws) == 2 else orig(*a, **k)

    torch.randn = fake_randn
    try:
        with RH._scratch_cwd(), torch.no_grad():
            pred, feat = diffusion.ddim_sample_loop_for_training(model, (1, 2, 64, 64), noise=None, clip_denoised=False, model_kwargs=kwargs,
                                                                 eta=0.0, progress=True, denoised_fn=None, sampling_kwargs=None, logger=None,
                                                                 n_batch=1, time_variant=True, iter=True, mode="train", timestep=0, pyramid=None)
    finally:
        torch.randn = orig
    np.savez_compressed(os.path.join(OUT, "rollout_doc0_t0.npz"), pred=pred.numpy(), feat_sub=feat[:, ::16, ::4, ::4].numpy())
    print("rollout golden: pred std", float(pred.std()), "range", float(pred.min()), float(pred.max()))


if __name__ == "__main__":
    main()
