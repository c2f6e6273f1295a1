#!/usr/bin/env python
"""bench.py — dewarped docs/sec (DDIM sampling + unwarp) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3                  # our arm, N=1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W    # N>1 (document-sharded, no collective)
    python bench.py --impl reference ...                             # reference's CPU path (oracle port)

One "step" = one batch of `--docs` synthetic documents per GPU through the whole hot path:
static conditioning (pyramid, embeds) -> S-step DDIM sampling (n_batch hypotheses) -> hypothesis
mean -> fused upsample + bilinear unwarp of the H x W photo.  Workload at N=1 = BASELINE.json
configs[1] (val_TDiff batch 1, 2000x1500 photo, S=3, n_batch=2).
  value : docs/s with the inputs already resident in HBM
  e2e   : the same through the public API with pinned-host inputs (H2D inside the timed region)
          and the unwarped uint8 image read back to the host (D2H inside the timed region)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DVD_PRECISION", "bf16x3"), choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--docs", type=int, default=1, help="documents per step per GPU")
    ap.add_argument("--height", type=int, default=1500)
    ap.add_argument("--width", type=int, default=2000)
    ap.add_argument("--diffusion-steps", type=int, default=3)
    ap.add_argument("--n-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"val_TDiff batch {a.docs}/GPU: S={a.diffusion_steps} DDIM steps x n_batch={a.n_batch} hypotheses + unwarp of a "
            f"{a.width}x{a.height} (WxH) synthetic photo")


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            hi = [s for s in sm if s >= 0.5 * max(sm)] or sm          # samples under load
            out = {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def algorithmic_gflop_per_doc(S, n_batch):
    """live, hoisted FLOPs only."""
    return 101.87 + n_batch * S * 262.52


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_doc_seconds(a, sd, inp, photo, full: bool):
    """The reference's own CPU implementation of the path, as written (12 DiT blocks, nothing hoisted), via the oracle
    port.  A bounded sample: ONE of the S denoiser forwards (n_batch hypotheses) is executed and scaled by S, plus the
    full-resolution upsample+grid_sample+uint8 cast.  Returns seconds per document."""
    from oracle import dvd_oracle as O
    n = a.n_batch
    rep = lambda v: v.repeat(n, 1, 1, 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        if full:
            m = O.sample(sd, inp, S=a.diffusion_steps, n_batch=n, as_written=True)
            t_den = time.perf_counter() - t0
        else:
            sch = O.Schedule(a.diffusion_steps)
            pred, _ = O.denoiser_forward(sd, inp["x_T"], sch.scaled_t(a.diffusion_steps - 1), rep(inp["init_flow"]), rep(inp["init_feat"]),
                                         None, y512=rep(inp["y512"]), mask_cat=rep(inp["mask_cat"]), mask_y512=rep(inp["mask_y512"]),
                                         line_msk=rep(inp["line_msk"]), as_written=True)
            t_den = (time.perf_counter() - t0) * a.diffusion_steps
            m = torch.clamp(pred.mean(0, keepdim=True), -1, 1)
        t1 = time.perf_counter()
        img = O.unwarp(m, photo)
        _ = O.to_uint8_hwc(img)
        t_unw = time.perf_counter() - t1
    return t_den + t_unw, t_den, t_unw


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth_workload as synth
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(1234)
    inp = synth.make_doc_inputs(0, H=a.height, W=a.width)
    photo = inp.pop("photo")
    for _ in range(min(a.warmup, 1)):
        cpu_reference_doc_seconds(a, sd, inp, photo, full=False)
    ts = [cpu_reference_doc_seconds(a, sd, inp, photo, full=False)[0] for _ in range(a.steps)]
    sec = sum(ts) / len(ts)
    value = 1.0 / sec
    line = {"impl": "reference", "metric": "dewarped docs/sec (sampling+unwarp)", "value": value, "unit": "docs/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": "docs/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "oracle port of the reference as written (12 DiT blocks, no hoisting); per step: 1 of the "
                                       f"{a.diffusion_steps} denoiser forwards x{a.diffusion_steps} + full-size unwarp"},
            "e2e": {"value": value, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch.distributed as dist
    import synth_workload as synth                          # synthetic workload generator (no oracle code on this arm)
    import dvd_b200
    from dvd_b200 import _lib
    from dvd_b200.model import DiT
    from dvd_b200.pipeline import DewarpPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    # ---- synthetic workload: documents sharded by rank (doc id = step*world*docs + rank*docs + j); no collective on the path
    sd = synth.make_state_dict(1234, live_only=True)
    model = DiT(precision=a.precision)
    model.load_state_dict(sd, strict=False)
    model.to(dev)
    pipe = DewarpPipeline(model, diffusion_steps=a.diffusion_steps, n_batch=a.n_batch, docs=a.docs, height=a.height, width=a.width)
    n_var = 2                                               # distinct input sets that the steps rotate through
    host_sets = []
    for v in range(n_var):
        docs = [synth.make_doc_inputs(1000 * rank + v * a.docs + j, H=a.height, W=a.width) for j in range(a.docs)]
        hs = {k: torch.cat([d[k] for d in docs]).contiguous().pin_memory() for k in ("y512", "mask_cat", "mask_y512", "line_msk", "x_T")}
        hs["photo_u8"] = torch.cat([d["photo"] for d in docs]).permute(0, 2, 3, 1).to(torch.uint8).contiguous().pin_memory()
        host_sets.append(hs)
    dev_sets = [{k: v.to(dev) for k, v in hs.items()} for hs in host_sets]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        pipe.kernel_launches = 0
        t0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)                                             # L2 flush between timed iterations (not timed)
            evs[i][0].record()
            fn(warmup + i)
            evs[i][1].record()
        barrier()
        wall = time.perf_counter() - t0
        launches = pipe.kernel_launches
        ms = [s.elapsed_time(e) for s, e in evs]
        return ms, wall, launches

    # (1) device-resident inputs
    out_dev = [None]

    def step_dev(i):
        out_dev[0] = pipe.run_device(dev_sets[i % n_var])

    # (2) end to end through the public API: pinned host -> device -> ... -> host uint8 image
    def step_e2e(i):
        pipe.run_host(host_sets[i % n_var])

    def timed_e2e_pipelined(steps, warmup):
        """Throughput form of the same API (submit_host / wait, two batches outstanding): the uploads of batch i+1 and the download of
        batch i overlap the kernels in between.  Every batch's H2D and D2H copies are inside the timed region; no L2 flush here (each
        step streams 20+ MB of new inputs and > 126 MB of weights)."""
        pending = None
        for i in range(warmup):
            t = pipe.submit_host(host_sets[i % n_var])
            if pending is not None:
                pipe.wait(pending)
            pending = t
        pipe.wait(pending); pending = None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            t = pipe.submit_host(host_sets[(warmup + i) % n_var])
            if pending is not None:
                pipe.wait(pending)
            pending = t
        pipe.wait(pending)                                                    # the last image is on the host
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.05)
    ms_dev, wall_dev, launches = timed(step_dev, a.steps, a.warmup)
    ms_e2e, wall_e2e, _ = timed(step_e2e, a.steps, a.warmup)              # synchronous calls: per-document latency
    ms_e2e_pipe = timed_e2e_pipelined(a.steps, a.warmup)                     # two batches outstanding: throughput
    clocks = sampler.stop() if sampler else None

    tot_dev, tot_e2e = sum(ms_dev) / 1e3, min(sum(ms_e2e), ms_e2e_pipe) / 1e3
    if world > 1:
        t = torch.tensor([tot_dev, tot_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                              # max over ranks (timing only; not on the data path)
        tot_dev, tot_e2e = float(t[0]), float(t[1])
    docs_total = a.docs * a.steps * world
    value, e2e = docs_total / tot_dev, docs_total / tot_e2e

    if rank == 0:
        peaks, which = measured_peaks()
        prof = pipe.profile_kernels(dev_sets[0])                             # CUDA-event timing of the dominant kernels, L2-flushed
        gflop = algorithmic_gflop_per_doc(a.diffusion_steps, a.n_batch) * a.docs
        line = {"metric": "dewarped docs/sec (sampling+unwarp)", "value": value, "unit": "docs/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": tot_dev / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
                "config": {"workload": workload_name(a), "docs_per_step_per_gpu": a.docs, "photo_hw": [a.height, a.width],
                           "diffusion_steps": a.diffusion_steps, "n_batch": a.n_batch, "weights": "random-init (synth_workload.py seed 1234)",
                           "l2": "256 MiB flush write between timed iterations (value, synchronous e2e); pipelined e2e: no flush, every step streams new inputs and > 126 MB of weights", "cuda_graph": pipe.use_graph, "parallelism": f"document-sharded x{world}, no collective"},
                "p50_latency_ms": statistics.median(ms_dev),
                "e2e": {"value": e2e, "unit": "docs/s", "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                        "p50_latency_ms": statistics.median(ms_e2e), "io": "uint8 HWC photo in, uint8 HWC dewarped image out",
                        "mode": "submit_host/wait, two batches outstanding (uploads and downloads overlap the kernels of the batch in between)",
                        "synchronous_docs_per_s": a.docs * a.steps * world / (sum(ms_e2e) / 1e3)},      # rank 0's clock x world
                "gpu_launches": int(launches),
                "clocks": clocks,
                "denoiser_tflops_effective": gflop / 1e3 / (tot_dev / a.steps / 1.0) if tot_dev > 0 else None,
                "roofline": prof["roofline"], "roofline_unwarp": prof["roofline_unwarp"], "kernel_share": prof.get("share"),
                "peaks": which}
        if not a.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count())
            sd_full = synth.make_state_dict(1234)
            inp = synth.make_doc_inputs(0, H=a.height, W=a.width)
            photo = inp.pop("photo")
            sec, t_den, t_unw = cpu_reference_doc_seconds(a, sd_full, inp, photo, full=False)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "docs/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"1 document: 1 of {a.diffusion_steps} as-written denoiser forwards x{a.diffusion_steps} "
                                              f"({t_den:.1f} s) + full-size unwarp ({t_unw:.2f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
