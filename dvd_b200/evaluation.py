"""Drop-in replacement for the per-document evaluation driver
(train_settings/dvd/evaluation.py:142-327 ``run_evaluation_docunet``), same signature.

What stays the reference's: the preprocessing networks passed in (``pretrained_dewarp_model`` = GeoTr_Seg_Inf,
``pretrained_seg_model`` = Seg/U2NETP, ``pretrained_line_seg_model`` = UNet) and the dataset/loader.  What is replaced:
the sampler call (evaluation.py:80-138), the upsample + base + affine (:300-306) and the unwarp + uint8 conversion
(visualization_utils.py:64-78), which run in libdvd_b200.  Extensions: documents are sharded over the ranks of an
initialised process group (``dvd_b200.dist``), per-document device timings are gathered at the end; the full-resolution
photo is uploaded as uint8 HWC (a quarter of the fp32 bytes; the loader's float photo holds integers 0..255,
doc_benchmark.py:68-81) and the PNG encode (visualization_utils.py:76-78) runs on worker threads off the GPU's critical path.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import dist as D
from .unwarp import dewarp_fullres


def run_sample_lr_dewarping(settings, logger, diffusion, model, radius, source, feature_size, raw_corr, init_flow, c20, source_64,
                            pyramid, doc_mask, seg_map_all=None, textline_map=None, init_feat=None):
    """evaluation.py:80-138 (same positional surface)."""
    kw = {"init_flow": init_flow, "src_feat": c20, "src_64": None, "y512": source, "tmode": settings.env.train_mode,
          "mask_cat": doc_mask, "init_feat": init_feat, "iter": settings.env.iter}
    if settings.env.use_gt_mask is False:
        kw["mask_y512"] = seg_map_all
    if settings.env.use_line_mask is True:
        kw["line_msk"] = textline_map
    if logger is not None:
        logger.info("\nStarting sampling")
    sample, _ = diffusion.ddim_sample_loop(model, (source.shape[0], 2, feature_size, feature_size), noise=None,
                                           clip_denoised=settings.env.clip_denoised, model_kwargs=kw, eta=0.0, progress=False,
                                           denoised_fn=None, sampling_kwargs={"src_img": source}, logger=logger,
                                           n_batch=settings.env.n_batch, time_variant=settings.env.time_variant, pyramid=pyramid)
    return sample            # already clamped to [-1, 1] by the hypothesis-mean kernel (gaussian_diffusion.py:640, evaluation.py:137)


def save_dewarped(settings, image_u8_hwc: np.ndarray, data_path, root: str = "."):
    """visualization_utils.py:64-78 file layout (`root` = the working directory at call time: the encode may run on a worker thread)."""
    from PIL import Image
    d = os.path.join(root, f"vis_hp/{settings.env.eval_dataset_name}/{settings.name}/dewarped_pred")
    os.makedirs(d, exist_ok=True)
    os.makedirs(os.path.join(root, f"vis_hp/{settings.env.eval_dataset_name}/{settings.name}/pred_flow"), exist_ok=True)
    name = data_path[0].split("/")[-1][:-4]
    Image.fromarray(image_u8_hwc).save(f"{d}/warped_{name}.png")


def photo_as_uint8_hwc(source_vis: torch.Tensor):
    """[B,3,H,W] float photo holding integers 0..255 (doc_benchmark.py:77-81: ArrayToTensor of a uint8 image) -> uint8 [B,H,W,3], or
    None when the values are not exactly representable (then the fp32 photo is uploaded like the reference does)."""
    if source_vis.dtype == torch.uint8:
        return source_vis.permute(0, 2, 3, 1).contiguous()
    u8 = source_vis.to(torch.uint8)
    if not torch.equal(u8.to(source_vis.dtype), source_vis):
        return None
    return u8.permute(0, 2, 3, 1).contiguous()


@torch.no_grad()
def run_evaluation_docunet(settings, logger, val_loader, diffusion, model, pretrained_dewarp_model, pretrained_line_seg_model=None,
                           pretrained_seg_model=None):
    os.makedirs(f"vis_hp/{settings.env.eval_dataset_name}/{settings.name}", exist_ok=True)
    dev = model.device
    # Documents are sharded over the ranks of an INITIALISED process group only (then gather_metrics really sees every rank's
    # timings); the loader must be the reference's un-sharded one (val_TDiff.py:104) - with a DistributedSampler, pass shard=False.
    import torch.distributed as tdist
    shard = getattr(settings.env, "shard_documents", True) and tdist.is_available() and tdist.is_initialized()
    rank = tdist.get_rank() if shard else 0
    world = tdist.get_world_size() if shard else 1
    image_size = 64
    times = {}
    from concurrent.futures import ThreadPoolExecutor
    pool, pending, cwd = ThreadPoolExecutor(max_workers=int(os.environ.get("DVD_PNG_THREADS", "4"))), [], os.getcwd()
    for i, data in enumerate(val_loader):
        if i % world != rank:                                                    # document sharding (no collective)
            continue
        data_path = data["path"]
        source_288 = F.interpolate(data["source_image"], size=288, mode="bilinear", align_corners=True).to(dev)     # evaluation.py:162
        B = data["source_image"].shape[0]
        init_feat = torch.zeros((B, 256, image_size, image_size), dtype=torch.float32, device=dev)                # :166-169
        ref_bm, mask_x = pretrained_dewarp_model(source_288)                                                       # :172-173
        if settings.env.use_init_flow:
            init_flow = F.interpolate(ref_bm / 287.0, size=image_size, mode="bilinear", align_corners=True)       # :176-178
        else:
            init_flow = torch.zeros((B, 2, image_size, image_size), dtype=torch.float32, device=dev)              # :180
        source = data["source_image"].to(dev)                                                                     # prepare_data :33
        source_vis = data["source_image_ori"] if "source_image_ori" in data else data["source_image"]             # :21-24
        mskx, d0, hx6, hx5d, hx4d, hx3d, hx2d, hx1d = pretrained_seg_model(source_288)                             # :204
        up = lambda v: F.interpolate(v, size=image_size, mode="bilinear", align_corners=False)
        seg_map_all = torch.cat((up(hx6), up(hx5d), up(hx4d), up(hx3d), up(hx2d), up(hx1d)), dim=1)               # :205-212
        textline_map, _ = pretrained_line_seg_model(mskx)                                                          # :215
        textline_map = up(textline_map)                                                                           # :216
        torch.cuda.synchronize(dev)
        t0 = time.time()
        sample = run_sample_lr_dewarping(settings, logger, diffusion, model, 4, source, image_size, None, init_flow, None, None, None,
                                         mask_x, seg_map_all, textline_map, init_feat)                            # :247-264
        torch.cuda.synchronize(dev)
        times[i] = time.time() - t0
        if settings.env.visualize:
            u8 = photo_as_uint8_hwc(source_vis)
            if u8 is not None:
                img = dewarp_fullres(sample, u8.to(dev, non_blocking=True))                                        # :300-306 + :317-318, 6 B/px
            else:
                img = dewarp_fullres(sample, source_vis.to(dev).float(), out_uint8=True)
            host = img[0].cpu().numpy()
            pending.append(pool.submit(save_dewarped, settings, host, data_path, cwd))                             # PNG encode off the critical path
    for fut in pending:
        fut.result()
    pool.shutdown()
    allt = {}
    for d in D.gather_metrics(times):
        allt.update(d)
    if rank == 0 and allt:
        print(len(allt))
        print("Elapsed time:{:.4f} avg_second ".format(sum(allt.values()) / len(allt)))                            # evaluation.py:326-327
    return allt
