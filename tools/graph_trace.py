"""In-graph kernel timeline of one document (bf16, batch 1) from CUPTI via torch.profiler.

ncu durations are cold-cache and serialised; this shows what the kernels cost INSIDE the CUDA-graph replay (with programmatic
dependent launch overlap): per-kernel duration summed by name and by (name, grid), plus the idle gaps between consecutive kernels.
Usage: python tools/graph_trace.py [--docs 1] [--out gpurun_out/graph_trace.txt]"""
import argparse, collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_workload as synth                             # synthetic workload generator
from dvd_b200.model import DiT
from dvd_b200.pipeline import DewarpPipeline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1)
    ap.add_argument("--height", type=int, default=1500)
    ap.add_argument("--width", type=int, default=2000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = DiT(precision=os.environ.get("DVD_PRECISION", "bf16x3"))
    model.load_state_dict(synth.make_state_dict(1234, live_only=True), strict=False)
    model.to(dev)
    pipe = DewarpPipeline(model, diffusion_steps=3, n_batch=2, docs=a.docs, height=a.height, width=a.width)
    docs = [synth.make_doc_inputs(j, H=a.height, W=a.width) for j in range(a.docs)]
    hs = {k: torch.cat([d[k] for d in docs]).contiguous() for k in ("y512", "mask_cat", "mask_y512", "line_msk", "x_T")}
    hs["photo_u8"] = torch.cat([d["photo"] for d in docs]).permute(0, 2, 3, 1).to(torch.uint8).contiguous()
    ds = {k: v.to(dev) for k, v in hs.items()}
    for _ in range(4):
        pipe.run_device(ds)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            pipe.run_device(ds)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name and "emset" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    # last replay = after the last unwarp-of-the-previous-run
    idx = [i for i, e in enumerate(evs) if "k_unwarp" in e.name]
    run = evs[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else evs
    lines = []
    t0, t1 = run[0].time_range.start, run[-1].time_range.end
    busy = sum(e.time_range.end - e.time_range.start for e in run)
    lines.append(f"# one document inside the graph replay: {len(run)} kernels, span {t1 - t0:.1f} us, summed kernel time {busy:.1f} us "
                 f"(overlap/gaps: {t1 - t0 - busy:+.1f} us)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    gaps = collections.defaultdict(lambda: [0, 0.0])
    prev_end = None
    for e in run:
        name = e.name.replace("void ", "").replace("dvd::", "").split("(")[0]
        d = e.time_range.end - e.time_range.start
        agg[name][0] += 1; agg[name][1] += d
        if prev_end is not None:
            g = e.time_range.start - prev_end
            gaps[name][0] += 1; gaps[name][1] += g
        prev_end = max(prev_end or 0, e.time_range.end)
    lines.append(f"{'kernel':52s} {'n':>4s} {'total_us':>10s} {'avg_us':>8s} {'share':>7s} {'gap_before_avg_us':>18s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        g = gaps.get(k, [1, 0.0])
        lines.append(f"{k[:52]:52s} {v[0]:4d} {v[1]:10.1f} {v[1] / v[0]:8.1f} {100 * v[1] / busy:6.1f}% {g[1] / max(g[0], 1):18.2f}")
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")
        # full sequence for one decoder layer's worth of inspection
        with open(a.out.replace(".txt", "_seq.txt"), "w") as f:
            for e in run:
                f.write(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:110]}\n")


if __name__ == "__main__":
    main()
