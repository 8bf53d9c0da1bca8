"""Times fssb200_eval_levelmajor (device arrays) for C2 / C3 / C5-HT shapes; run once per FSSB200_LM_MODE (2 = direct
loads, default 7 = TMA tiles) -- the mode is read once per process.  Prints evals/s and the LDS-ceiling fraction
(9.22 T lookups/s, 160 per AES block)."""
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

    for scheme, n, group, k, blocks in (("dpf", 32, "bytes", 1 << 22, 32), ("dcf", 64, "u128", 1 << 21, 128),
                                        ("halftree", 32, "bytes", 1 << 22, 32), ("dpf", 128, "u128", 1 << 20, 128)):
        ctx = fss_b200.Context(scheme, n, group)
        s0s, betas = rnd((k, 2, 4)), rnd((k, 4))
        s0s[..., 3] &= ~1
        betas[..., 3] &= ~1
        alphas = torch.randint(0, 1 << min(n, 62), (k,), dtype=torch.int64, device=dev, generator=g)
        xs = torch.randint(0, 1 << min(n, 62), (k,), dtype=torch.int64, device=dev, generator=g)
        if n <= 32:
            alphas, xs = alphas.to(torch.int32), xs.to(torch.int32)
        r = ctx.gen(s0s, alphas, betas)
        cws, ocws = r if scheme == "halftree" else (r, None)
        seeds = s0s[:, 0].contiguous()
        want = ctx.eval(0, seeds, cws, xs, ocws)
        lay = ctx.relayout(cws)
        del cws
        ys = torch.empty((k, 4), dtype=torch.int32, device=dev)
        for _ in range(3):
            ctx.eval_levelmajor(0, seeds, lay, xs, ocws, out=ys)
        assert torch.equal(ys, want), (scheme, n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ctx.eval_levelmajor(0, seeds, lay, xs, ocws, out=ys)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"LM_MODE={os.environ.get('FSSB200_LM_MODE', '7')} {scheme} n={n} keys=2^{k.bit_length() - 1}: {ms:.3f} ms  "
              f"{k / ms / 1e6:.3f} G evals/s  lds_frac {k * blocks * 160 / (ms * 1e-3) / 9.22e12:.4f}", flush=True)
        del lay, ys, want


if __name__ == "__main__":
    main()
