"""Micro-benchmark of the tcgen05 GEMM on the denoiser's shapes (CUDA events, L2 flushed between runs).
Usage: python tools/gemm_bench.py            (env DVD_GEMM_V1=1 selects the non-persistent kernel, DVD_GEMM_BN forces a tile width)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvd_b200 import _lib

SHAPES = [(2048, 4608, 1536, "dec qkv"), (2048, 1536, 1536, "dec fc"), (2048, 2048, 1536, "dec conv1"), (2048, 1536, 2048, "dec conv2"),
          (8192, 1152, 384, "dit qkv"), (8192, 1536, 384, "dit fc1"), (8192, 384, 1536, "dit fc2"), (8192, 384, 384, "dit proj"),
          (16384, 4608, 1536, "dec qkv x8 docs"), (16384, 1536, 1536, "dec fc x8 docs")]


def main():
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tag = f"v1={os.environ.get('DVD_GEMM_V1','0')} bn={os.environ.get('DVD_GEMM_BN','auto')}"
    for M, N, K, name in SHAPES:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        st = _lib.stream_ptr()
        run = lambda: _lib.check(lib.dvd_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(b), _lib.ptr(out), None, M, N, K, st), "gemm")
        for _ in range(3):
            run()
        ts = []
        for i in range(10):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        t = ts[len(ts) // 2]
        ref = (A.float() @ W.float().t() + b)
        err = float((out.float() - ref).abs().max() / ref.abs().max())
        print(f"{tag:18s} {name:18s} M={M:6d} N={N:5d} K={K:5d}  {t:7.1f} us  {2.0 * M * N * K / t / 1e6:7.1f} TF/s  relerr {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
