"""Runs the UNMODIFIED reference for the hot path (TEST / BASELINE INFRASTRUCTURE, never the product path).

Where the reference comes from: ``/root/reference`` in the build container; on the GPU box that tree does not exist, so
``__graft_entry__.build()`` copies it verbatim (Python sources only, no edits) to the git-ignored ``baseline/_ref/`` which
travels with the gpurun snapshot.  Users: ``oracle/make_golden.py`` (golden fixtures), ``tests/test_reference_dropin.py``
and the baseline legs of ``bench.py`` (``--impl reference``, ``cpu_baseline``, ``torch_b200``, ``preprocessing``).

Import shims (``oracle/ref_shims``) restate the few third-party classes the reference imports but
this image lacks (timm, mmcv, mmengine, mpi4py, blobfile, matplotlib, h5py); the reference code
itself is imported as is.
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "ref_shims")
BASELINE_REF = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def _find_root() -> str:
    for cand in (os.environ.get("DVD_REFERENCE_ROOT"), BASELINE_REF, "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "train_settings", "dvd")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def install_baseline_ref(src: str = "/root/reference") -> bool:
    """Verbatim copy of the reference's Python sources to baseline/_ref (git-ignored, ships with gpurun).  The reference has no
    setup.py / pyproject, so there is nothing to pip-install: it is imported from this directory through sys.path."""
    import shutil
    if not os.path.isdir(os.path.join(src, "train_settings", "dvd")):
        return False
    if os.path.isdir(BASELINE_REF):
        shutil.rmtree(BASELINE_REF)
    shutil.copytree(src, BASELINE_REF, ignore=shutil.ignore_patterns(".git", "asset", "matlab_code", "__pycache__", "*.pyc", "*.png", "*.jpg"))
    return True


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "train_settings", "dvd"))


def _setup_path():
    # the reference's top-level package is called `datasets`: keep it FIRST on sys.path
    for p in (_SHIMS, REF_ROOT):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


@contextlib.contextmanager
def _scratch_cwd():
    """The reference writes debug PNGs to ./vis_hp/debug_vis (gaussian_diffusion.py:606,614)."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "vis_hp", "debug_vis"))
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


def build_reference_model(sd):
    _setup_path()
    from train_settings.dvd.improved_diffusion.cross_model import DiT_models2
    model = DiT_models2["DiT-S/2"](input_size=64, in_channels=2, tv=True)   # script_util.py:155-162
    model.load_state_dict(sd, strict=True)
    return model.eval()


def build_reference_diffusion(S: int = 3, schedule: str = "cosine"):
    _setup_path()
    from train_settings.dvd.improved_diffusion import gaussian_diffusion as gd
    from train_settings.dvd.improved_diffusion.respace import SpacedDiffusion, space_timesteps
    betas = gd.get_named_beta_schedule(schedule, S)
    return SpacedDiffusion(                                                  # script_util.py:206-244 with local.py values
        use_timesteps=space_timesteps(S, [S]), betas=betas,
        model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_LARGE,
        loss_type=gd.LossType.RESCALED_MSE, rescale_timesteps=True)


class Recorder(torch.nn.Module):
    """Transparent wrapper that records what the sampler feeds the denoiser and what it returns."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.calls = []

    def forward(self, x, t, **kw):
        out, feat = self.model(x, t, **kw)
        self.calls.append({"x": x.detach().clone(), "t": t.detach().clone(), "init_flow": kw["init_flow"].detach().clone(),
                           "init_feat": kw["init_feat"].detach().clone(), "pred": out.detach().clone(), "feat": feat.detach().clone()})
        return out, feat


def reference_sample(model, inp: dict, S: int = 3, n_batch: int = 2, seed: int | None = None, schedule: str = "cosine"):
    """evaluation.py:80-138 run_sample_lr_dewarping's sampler call with the default settings of
    admin/local.py.  Returns (sample[1,2,64,64] clamped, recorder)."""
    diffusion = build_reference_diffusion(S, schedule)
    rec = Recorder(model)
    kwargs = {"init_flow": inp["init_flow"].clone(), "src_feat": None, "src_64": None, "y512": inp["y512"],
              "tmode": "stage_1_dit_cross", "mask_cat": inp["mask_cat"], "init_feat": inp["init_feat"].clone(),
              "iter": True, "mask_y512": inp["mask_y512"], "line_msk": inp["line_msk"]}
    if seed is not None:
        torch.manual_seed(seed)
    with _scratch_cwd(), torch.no_grad():
        sample, _ = diffusion.ddim_sample_loop(
            rec, (1, 2, 64, 64), noise=None, clip_denoised=False, model_kwargs=kwargs, eta=0.0, progress=False,
            denoised_fn=None, sampling_kwargs={"src_img": inp["y512"]}, logger=None, n_batch=n_batch,
            time_variant=True, pyramid=None)
    return torch.clamp(sample, min=-1, max=1), rec                            # evaluation.py:137


def reference_unwarp(map64: torch.Tensor, photo: torch.Tensor):
    """evaluation.py:300-306 (inline in the reference) + visualization_utils.py:75 via the
    reference's own coords_grid_tensor and register_model2.  Returns (grid[1,2,H,W], image[1,C,H,W])."""
    _setup_path()
    import torch.nn.functional as F
    from datasets.utils.warping import register_model2
    from train_settings.dvd.improved_diffusion.gaussian_diffusion import coords_grid_tensor
    H, W = photo.shape[-2:]
    sample = F.interpolate(map64, size=(H, W), mode="bilinear", align_corners=True)
    base = F.interpolate(coords_grid_tensor((512, 512)) / 511., size=(H, W), mode="bilinear", align_corners=True)
    sample = (((sample + base.to(sample.device)) * 1) * 2 - 1) * 0.987
    reg = register_model2((512, 512), "bilinear")
    return sample, reg([photo.float(), sample])


# ----------------------------------------------------------------------------------------------- baseline timing legs (bench.py)
def reference_document(model, inp: dict, photo: torch.Tensor, S: int, n_batch: int, seed: int, device: str = "cpu"):
    """One whole document through the reference's own code: evaluation.py:80-138 (sampling, incl. its three debug PNG dumps) +
    :300-306 (upsample / affine) + visualization_utils.py:75-77 (grid_sample, uint8 cast).  Returns the HWC uint8 image."""
    dev = torch.device(device)
    inp_d = {k: v.to(dev) for k, v in inp.items()}
    sample, _ = reference_sample(model, inp_d, S=S, n_batch=n_batch, seed=seed)
    _, img = reference_unwarp(sample, photo.to(dev))
    return img[0].permute(1, 2, 0).cpu().numpy().astype("uint8")


def build_preprocessing_nets():
    """The three preprocessing networks of val_TDiff.py:58-74, random-init (the checkpoints are not in the tree)."""
    _setup_path()
    from train_settings.models.geotr.geotr_core import GeoTr_Seg_Inf, Seg
    from train_settings.models.geotr.unet_model import UNet
    return {"GeoTr_Seg_Inf": GeoTr_Seg_Inf().eval(), "Seg(U2NETP)": Seg().eval(), "line UNet": UNet(n_channels=3, n_classes=1).eval()}
