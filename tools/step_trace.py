"""Phase timeline of ONE GEMM of a real denoiser step (instrumented build, launches filtered by N at run time):
    DVD_NVCC_EXTRA="-DDVD_GEMM_TRACE -DDVD_GEMM_TRACE2" python -m dvd_b200.build
    DVD_LIB=dvd_b200/libdvd_b200_trace.so DVD_NO_GRAPH=1 python tools/step_trace.py [N ...]
Runs documents eagerly and prints the trace of the LAST launch with that N (2048: conv1 of decoder layer 5 in the last step;
1536: conv2; 4608: q|k|v)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_workload as synth
from dvd_b200 import _lib
from dvd_b200.model import DiT
from dvd_b200.pipeline import DewarpPipeline
from gemm_trace import NAMES


def main():
    dev = torch.device("cuda", 0)
    model = DiT(precision="bf16x3")
    model.load_state_dict(synth.make_state_dict(1234, live_only=True), strict=False)
    model.to(dev)
    pipe = DewarpPipeline(model, diffusion_steps=3, n_batch=2, docs=1, height=1500, width=2000)
    d = synth.make_doc_inputs(0, H=1500, W=2000)
    ds = {k: d[k].to(dev).contiguous() for k in ("y512", "mask_cat", "mask_y512", "line_msk", "x_T")}
    ds["photo_u8"] = d["photo"].permute(0, 2, 3, 1).to(torch.uint8).contiguous().to(dev)
    for _ in range(2):
        pipe.run_device(ds)
    torch.cuda.synchronize()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for N in [int(a) for a in sys.argv[1:]] or [2048]:
        raw.dvd_debug_pair_trace_filter(N)
        for _ in range(2):
            pipe.run_device(ds)
        torch.cuda.synchronize()
        print(f"== N = {N}")
        dump(raw)


def dump(raw):
    n = 512
    buf = (ctypes.c_ulonglong * (n * 16))()
    f = raw.dvd_debug_pair_trace; f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert f(buf, n) == 0
    t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 16).astype(np.int64)
    t = t[t[:, 0] > 0]
    t = t[t[:, 0] > t[:, 0].max() - 1_000_000]
    t0 = t[:, 0].min()
    names = list(NAMES)
    names[7:13] = ["e: tmem ld done", "e: staged+sync", "e: lds done", "e: loads issued", "e: stores issued", "e: chunk end"]
    print(f"last traced launch: {len(t)} CTAs")
    for i, nm in enumerate(names):
        col = t[:, i]
        col = col[col >= t0]
        if nm == "-" or len(col) == 0:
            continue
        r = (col - t0) / 1e3
        print(f"    {nm:16s} n={len(col):4d}  min {r.min():7.2f}  median {np.median(r):7.2f}  max {r.max():7.2f} us")


if __name__ == "__main__":
    main()
