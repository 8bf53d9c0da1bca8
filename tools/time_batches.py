"""Times fssb200_eval (DPF n=32, key-major Cw, device arrays) on batches from 2^13 to 2^22 keys: how evenly a batch that
does not fill every warp, and the last partial round of tiles, spread over the SMs.  Prints ms, evals/s, LDS fraction."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fss_b200


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(2)

    def rnd(shape):
        return torch.randint(-2**31, 2**31 - 1, shape, dtype=torch.int64, device=dev, generator=g).to(torch.int32)

    for scheme in ("dpf", "halftree"):
        ctx = fss_b200.Context(scheme, 32, "bytes")
        k = 1 << 22
        s0s, betas, alphas, xs = rnd((k, 2, 4)), rnd((k, 4)), rnd((k,)), rnd((k,))
        s0s[..., 3] &= ~1
        betas[..., 3] &= ~1
        r = ctx.gen(s0s, alphas, betas)
        cws, ocws = r if scheme == "halftree" else (r, None)
        seeds = s0s[:, 0].contiguous()
        ys = torch.empty((k, 4), dtype=torch.int32, device=dev)
        for bits in (13, 14, 16, 18, 19, 20, 21, 22):
            kk = 1 << bits
            oc = None if ocws is None else ocws[:kk]
            for _ in range(3):
                ctx.eval(0, seeds[:kk], cws[:kk], xs[:kk], oc, out=ys[:kk])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                ctx.eval(0, seeds[:kk], cws[:kk], xs[:kk], oc, out=ys[:kk])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"{scheme} 2^{bits} keys: {ms * 1e3:.1f} us  {kk / ms / 1e6:.3f} G evals/s  lds_frac {kk * 32 * 160 / (ms * 1e-3) / 9.22e12:.4f}", flush=True)
        del cws, ys


if __name__ == "__main__":
    main()
