"""Reference-equivalent PyTorch ops on the GPU (the oracle port of the reference AS WRITTEN: 12 DiT blocks, nothing hoisted,
stock cuDNN/cuBLAS kernels) timed with CUDA events.  This is the denominator of north_star's ">= 10x the reference
PyTorch-on-B200 latency" target; it is a measurement aid (oracle code), not part of the product path.
    python tools/torch_gpu_baseline.py [H W]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dvd_oracle as O, synth


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1500, 2000)
    dev = torch.device("cuda:0")
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(1234).items()}
    inp = {k: v.to(dev) for k, v in synth.make_doc_inputs(0, H=H, W=W).items()}
    photo = inp.pop("photo")
    torch.backends.cudnn.benchmark = True                       # run_sampling.py:27
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        for hoist in (False, True):
            ts = []
            for it in range(5):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                with torch.no_grad():
                    e0.record()
                    m = O.sample(sd, inp, S=3, n_batch=2, as_written=not hoist)
                    e1.record()
                    img = O.unwarp(m, photo)
                    u8 = img[0].permute(1, 2, 0).to(torch.uint8)
                    e2.record()
                e2.synchronize()
                ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
            ts = ts[2:]
            s = sum(t[0] for t in ts) / len(ts); u = sum(t[1] for t in ts) / len(ts)
            print(f"torch-on-B200 tf32={tf32!s:5s} {'hoisted/live-only' if hoist else 'as written      '}: sampling {s:8.2f} ms  unwarp {u:6.3f} ms  "
                  f"-> {1000.0 / (s + u):7.2f} docs/s", flush=True)


if __name__ == "__main__":
    main()
