"""dvd_b200 — B200-native (sm_100a) implementation of DvD's sampling + unwarp hot path.

Public surface mirrors the reference modules it replaces:
    script_util.create_model_and_diffusion, model.DiT, sampler.SpacedDiffusion.ddim_sample_loop,
    unwarp.register_model2 / dewarp_fullres, evaluation.run_evaluation_docunet.
"""
from .model import DiT, DiT_models2                                   # noqa: F401
from .sampler import SpacedDiffusion, create_gaussian_diffusion       # noqa: F401
from .script_util import create_model_and_diffusion                   # noqa: F401
from .unwarp import dewarp_fullres, fullres_grid, register_model2     # noqa: F401

__version__ = "0.1.0"
