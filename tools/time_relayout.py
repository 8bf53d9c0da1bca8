"""Times fssb200_relayout alone (outputs preallocated) on a few shapes; prints keys/s and HBM bytes/s moved.

  python tools/time_relayout.py            (under gpurun)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fss_b200


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    for scheme, n, group, k in (("dpf", 32, "bytes", 1 << 22), ("dcf", 64, "u128", 1 << 21), ("halftree", 32, "bytes", 1 << 22),
                                ("dpf", 128, "u128", 1 << 20), ("dpf", 8, "u64", 1 << 22), ("dpf", 32, "bytes", 1 << 16)):
        ctx = fss_b200.Context(scheme, n, group)
        ncw = n if scheme == "halftree" else n + 1
        cws = torch.randint(-2**31, 2**31 - 1, (k, ncw, 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        lay = ctx.relayout(cws)
        for _ in range(3):
            ctx.relayout(cws, out=lay)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ctx.relayout(cws, out=lay)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        moved = cws.numel() * 4 + sum(t.numel() * 4 for t in lay if t is not None)
        print(f"{scheme} n={n} keys=2^{k.bit_length() - 1}: {ms:.3f} ms  {k / ms / 1e6:.3f} G keys/s  {moved / ms / 1e9:.2f} TB/s moved")
        del cws, lay


if __name__ == "__main__":
    main()
