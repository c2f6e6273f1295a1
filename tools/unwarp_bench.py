"""Standalone timing of the fused upsample+unwarp kernel (CUDA events, L2 flush between launches)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from dvd_b200 import dewarp_fullres


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    m = synth.make_map64(5, "smooth").to(dev)
    for (H, W) in [(1500, 2000), (4032, 3024)]:
        photo = synth.make_photo(H, W, 21, "page").to(dev)
        pu8 = photo[0].permute(1, 2, 0).to(torch.uint8).unsqueeze(0).contiguous()
        outf = torch.empty_like(photo); outu = torch.empty_like(pu8)
        for name, fn, bpp in (("f32->f32", lambda: dewarp_fullres(m, photo, out=outf), 24), ("u8->u8", lambda: dewarp_fullres(m, pu8, out=outu), 6)):
            ts = []
            for i in range(8):
                flush.fill_(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            t = sorted(ts[2:])[len(ts[2:]) // 2]
            print(f"unwarp {name:9s} {W}x{H}: {t:7.1f} us  {bpp * H * W / t / 1e3:7.1f} GB/s ({bpp} B/px)", flush=True)


if __name__ == "__main__":
    main()
