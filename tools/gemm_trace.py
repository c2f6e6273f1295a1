"""Phase timeline of one tcgen05 GEMM launch (needs the instrumented build: DVD_NVCC_EXTRA=-DDVD_GEMM_TRACE python -m dvd_b200.build).

Per CTA, %globaltimer at: start, prologue done (after griddepcontrol.wait), first stage landed, last stage landed, accumulator
complete, epilogue done.  Prints the distribution relative to the earliest CTA start."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvd_b200 import _lib


def main():
    lib = _lib.lib()
    raw = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else lib
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [(2048, 1536, 1536, "dec fc"), (2048, 2048, 1536, "dec conv1"), (2048, 1536, 2048, "dec conv2"), (2048, 4608, 1536, "dec qkv")]
    for M, N, K, name in shapes:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        st = _lib.stream_ptr()
        run = lambda: _lib.check(lib.dvd_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(b), _lib.ptr(out), None, M, N, K, st), "gemm")
        for _ in range(3):
            run()
        for cold in (True, False):
            if cold:
                flush.fill_(1)
            torch.cuda.synchronize()
            run(); torch.cuda.synchronize()
            n = 2048
            buf = (ctypes.c_ulonglong * (n * 8))()
            f = raw.dvd_debug_gemm_trace; f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int]
            assert f(buf, n) == 0
            t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 8).astype(np.int64)
            t = t[t[:, 0] > 0]
            # CTAs of THIS launch: those whose start is within 1 ms of the latest start
            t = t[t[:, 0] > t[:, 0].max() - 1_000_000]
            t0 = t[:, 0].min()
            r = (t[:, :6] - t0) / 1e3
            names = ["start", "prologue", "first stage", "last stage", "acc done", "epilogue"]
            print(f"{name} M={M} N={N} K={K} {'cold L2' if cold else 'warm L2'}: {len(t)} CTAs")
            for i, nm in enumerate(names):
                print(f"    {nm:12s} min {r[:, i].min():7.2f}  median {np.median(r[:, i]):7.2f}  max {r[:, i].max():7.2f} us")
            d = (t[:, 3] - t[:, 2]) / 1e3
            print(f"    main loop (first->last stage landed) median {np.median(d):.2f} us;  epilogue median {np.median((t[:, 5] - t[:, 4]) / 1e3):.2f} us")


if __name__ == "__main__":
    main()
