"""Micro-benchmark of the tcgen05 GEMM on the denoiser's shapes, both tensor modes.
Usage: python tools/gemm_bench.py [--check] [--graph]
  default : one launch per measurement between CUDA events, L2 flushed in between (includes the launch gap of an eager launch)
  --graph : 16 launches captured in ONE CUDA graph over rotating weight / output buffers (> L2 in total), replayed: the time per
            launch a step of the pipeline sees (no host time between kernels)
  --check : fewer runs (used by the parity tests)
  --epi=conv1|conv2 : (with --graph) the decoder's epilogues instead of bias + bf16: BN scale/shift + ReLU -> bf16 pair; BN + ReLU + fp32 residual
env: DVD_GEMM_V1=1 selects the generic single-CTA kernel; DVD_GEMM_BN forces the tile width of the persistent CTA-pair kernel; DVD_GEMM_DEBUG=1 prints the chosen configuration per launch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvd_b200 import _lib

SHAPES = [(2048, 4608, 1536, "dec qkv"), (2048, 1536, 1536, "dec fc"), (2048, 2048, 1536, "dec conv1"), (2048, 1536, 2048, "dec conv2"),
          (8192, 1152, 384, "dit qkv"), (8192, 1536, 384, "dit fc1"), (8192, 384, 1536, "dit fc2"), (8192, 384, 384, "dit proj"),
          (16384, 4608, 1536, "dec qkv x8 docs"), (16384, 1536, 1536, "dec fc x8 docs")]


def main():
    check, graph = "--check" in sys.argv, "--graph" in sys.argv
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tag = f"v1={os.environ.get('DVD_GEMM_V1','0')} bn={os.environ.get('DVD_GEMM_BN','auto')}"
    modes = [m for m in ("bf16", "bf16x3", "a16w3") if ("--" + m) in sys.argv] or ["bf16", "bf16x3"]     # a16w3: fp16 A x fp16 weight pair (two passes)
    for mode in modes:
        for M, N, K, name in SHAPES:
            A = torch.randn(M, K, device=dev) * 0.5; W = torch.randn(N, K, device=dev) / K ** 0.5
            wt = torch.float16 if mode == "a16w3" else torch.bfloat16          # a16w3: everything IEEE fp16
            Ah, Wh = A.to(wt), W.to(wt)
            Al = (A - Ah.float()).bfloat16() if mode == "bf16x3" else None
            Wl = (W - Wh.float()).to(wt) if mode != "bf16" else None
            b = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev)
            st = _lib.stream_ptr()

            epi = ([a.split("=")[1] for a in sys.argv if a.startswith("--epi=")] or [""])[0]
            sc, sh = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)

            def run(Wh_=Wh, Wl_=Wl, out16=None, out32=out, out_lo=None, res=None):
                if epi == "conv1" and out16 is not None:
                    return _lib.check(lib.dvd_gemm_tune(_lib.ptr(Ah), _lib.ptr(Al), K, _lib.ptr(Wh_), _lib.ptr(Wl_), K, None, _lib.ptr(sc), _lib.ptr(sh), 1,
                                                        _lib.ptr(out16), _lib.ptr(out_lo), None, None, M, N, K, _lib.stream_ptr()), "gemm")
                if epi == "conv2" and res is not None:
                    return _lib.check(lib.dvd_gemm_tune(_lib.ptr(Ah), _lib.ptr(Al), K, _lib.ptr(Wh_), _lib.ptr(Wl_), K, None, _lib.ptr(sc), _lib.ptr(sh), 1,
                                                        None, None, _lib.ptr(res), _lib.ptr(res), M, N, K, _lib.stream_ptr()), "gemm")
                _lib.check(lib.dvd_gemm_bf16(_lib.ptr(Ah), _lib.ptr(Al), K, _lib.ptr(Wh_), _lib.ptr(Wl_), K, _lib.ptr(b), _lib.ptr(out16), _lib.ptr(out32),
                                             M, N, K, _lib.stream_ptr()), "gemm")
            for _ in range(2 if check else 3):
                run()
            torch.cuda.synchronize()
            if graph:
                # bf16 output like the pipeline's GEMMs; 8 weight copies (8 x 14 MB for the QKV shape) so that weights stream from HBM
                R = 16
                Ws = [(Wh.clone(), Wl.clone() if Wl is not None else None) for _ in range(8)]
                nout = 16 if epi else 4                       # --epi: 16 cold output sets (> L2 together with the weights)
                outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nout)]
                los = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nout)] if epi == "conv1" else [None] * nout
                ress = [torch.zeros(M, N, device=dev) for _ in range(nout)] if epi == "conv2" else [None] * nout
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(R):
                        run(Ws[i % 8][0], Ws[i % 8][1], outs[i % nout], None, los[i % nout], ress[i % nout])
                g.replay(); torch.cuda.synchronize()
                ts = []
                for i in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3 / R)
                run()
            else:
                ts = []
                for i in range(3 if check else 10):
                    flush.fill_(i)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run(); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            t = ts[len(ts) // 2]
            if mode == "bf16":
                ref = Ah.double() @ Wh.double().t() + b.double()
            elif mode == "a16w3":
                ref = Ah.double() @ W.double().t() + b.double()
            else:
                ref = A.double() @ W.double().t() + b.double()
            err = float((out.double() - ref).abs().max() / ref.abs().max())
            print(f"{tag:24s} {mode:7s} {name:18s} M={M:6d} N={N:5d} K={K:5d}  {t:7.1f} us  {2.0 * M * N * K / t / 1e6:7.1f} TF/s  relerr {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
