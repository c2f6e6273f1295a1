"""Attention kernels alone: CUPTI kernel durations (torch.profiler) of dvd_test_attention's k_attn_* launch, warm L2, back to back.

    python tools/attn_bench.py [--B 2] [--T 1024] [--d 256] [--iters 20]
    DVD_LIB=dvd_b200/libdvd_b200_trace.so python tools/attn_bench.py --trace      (build: DVD_NVCC_EXTRA=-DDVD_ATTN_TRACE python -m dvd_b200.build)

--trace prints the phase timeline of cluster 0 of k_attn_pair (per 128-key step: Q K^T issued, P seen by the issuer, P V issued;
softmax warp 2: S seen, S in registers, exponentials done, P stored + arrived)."""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvd_b200 import _lib


def run(dev, B, H, T, d, prec, iters):
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(B, T, H * d, generator=g).to(dev) for _ in range(3))
    o = torch.empty(B, T, H * d, device=dev)
    scratch = torch.empty(max(B * H * T * T * 4, B * T * H * d * 12) + 1024, dtype=torch.uint8, device=dev)
    call = lambda: _lib.check(_lib.lib().dvd_test_attention(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), B, H, T, d, d ** -0.5, prec,
                                                            _lib.ptr(scratch), scratch.numel(), _lib.stream_ptr()), "dvd_test_attention")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            call()
        torch.cuda.synchronize()
    ds = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and "k_attn" in e.name:
            ds.setdefault(e.name.split("(")[0].replace("void ", "").replace("dvd::", ""), []).append(e.time_range.end - e.time_range.start)
    flops = 4.0 * B * H * T * T * d
    for name, v in ds.items():
        v.sort()
        med = v[len(v) // 2]
        print(f"B={B} H={H} T={T} d={d} prec={prec} {name:28s} n={len(v)} median {med:7.1f} us  min {v[0]:7.1f} us  {flops / med / 1e6:7.1f} TF/s")


def trace():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_ulonglong * 256)()
    rc = lib.dvd_debug_attn_trace(buf)
    assert rc == 0, rc
    rows = [[buf[t * 16 + s] for s in range(16)] for t in range(16)]
    t0 = min(x for r in rows for x in r if x)
    names = ["QK issued", "P seen", "PV issued", "S seen", "S in regs", "exp done", "P arrived", "PV all done", "O stored"]
    print("step " + " ".join(f"{n:>11s}" for n in names) + "   (ns from the first stamp)")
    for t, r in enumerate(rows):
        if not any(r):
            continue
        print(f"{t:4d} " + " ".join(f"{(r[s] - t0) if r[s] else 0:11d}" for s in range(len(names))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--H", type=int, default=6)
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--d", type=int, default=0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--trace", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    for d in ([a.d] if a.d else [256, 64]):
        for prec in (1, 2):
            run(dev, a.B if d == 256 else 4 * a.B, a.H, a.T, d, prec, a.iters)
    if a.trace:
        run(dev, a.B, a.H, a.T, 256, 2, 1)
        trace()


if __name__ == "__main__":
    main()
