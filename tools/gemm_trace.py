"""Phase timeline of the persistent CTA-pair GEMM (needs the instrumented build:
    DVD_NVCC_EXTRA=-DDVD_GEMM_TRACE python -m dvd_b200.build   ->  dvd_b200/libdvd_b200_trace.so
    DVD_LIB=dvd_b200/libdvd_b200_trace.so python tools/gemm_trace.py [--x3]).

Per CTA, %globaltimer at: start, prologue done (after griddepcontrol.wait), and for its first two units: first / last stage landed
(MMA warp), accumulator complete, TMEM drained, epilogue done (epilogue warp 2); then "epilogue warps done" and exit.  Prints the
distribution over CTAs relative to the earliest CTA start, warm L2 (what a step of the pipeline sees) and cold."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvd_b200 import _lib

NAMES = ["start", "prologue", "u0 first stage", "u0 last stage", "u0 acc done", "u0 tmem drained", "u0 epilogue", "u1 first stage", "u1 last stage",
         "u1 acc done", "u1 tmem drained", "u1 epilogue", "-", "-", "epi warps done", "exit"]


def main():
    x3 = "--x3" in sys.argv
    if "--epi" in sys.argv:                 # build with -DDVD_GEMM_TRACE -DDVD_GEMM_TRACE2: slots 7..12 = epilogue detail of (unit 0, chunk 0, warp 2)
        NAMES[7:13] = ["e: tmem ld done", "e: staged+sync", "e: lds done", "e: loads issued", "e: stores issued", "e: chunk end"]
    lib = _lib.lib()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [(2048, 1536, 1536, "dec fc"), (2048, 2048, 1536, "dec conv1"), (2048, 4608, 1536, "dec qkv"), (8192, 384, 384, "dit proj")]
    for M, N, K, name in shapes:
        A = torch.randn(M, K, device=dev) * 0.5; W = torch.randn(N, K, device=dev) / K ** 0.5
        Ah, Wh = A.bfloat16(), W.bfloat16()
        Al = (A - Ah.float()).bfloat16() if x3 else None
        Wl = (W - Wh.float()).bfloat16() if x3 else None
        b = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        run = lambda: _lib.check(lib.dvd_gemm_bf16(_lib.ptr(Ah), _lib.ptr(Al), K, _lib.ptr(Wh), _lib.ptr(Wl), K, _lib.ptr(b), _lib.ptr(out), None, M, N, K,
                                                   _lib.stream_ptr()), "gemm")
        for _ in range(3):
            run()
        for cold in (False, True):
            if cold:
                flush.fill_(1)
            torch.cuda.synchronize()
            run(); torch.cuda.synchronize()
            n = 512
            buf = (ctypes.c_ulonglong * (n * 16))()
            f = raw.dvd_debug_pair_trace; f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int]
            assert f(buf, n) == 0
            t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 16).astype(np.int64)
            t = t[t[:, 0] > 0]
            t = t[t[:, 0] > t[:, 0].max() - 1_000_000]            # CTAs of THIS launch
            t0 = t[:, 0].min()
            print(f"{name} M={M} N={N} K={K} {'x3' if x3 else 'bf16'} {'cold L2' if cold else 'warm L2'}: {len(t)} CTAs")
            for i, nm in enumerate(NAMES):
                if nm == "-":
                    continue
                col = t[:, i]
                col = col[col >= t0]                               # slots this CTA wrote in THIS launch (leader-only / second unit may be absent)
                if len(col) == 0:
                    continue
                r = (col - t0) / 1e3
                print(f"    {nm:16s} n={len(col):4d}  min {r.min():7.2f}  median {np.median(r):7.2f}  max {r.max():7.2f} us")


if __name__ == "__main__":
    main()
