"""A/B of the lanes-per-key choice of vdpf_finish_kernel (FSSB200_VDPF_FINISH_LANES) on Vdpf::EvalAll.

  python tools/ab_vdpf_finish.py           (under gpurun; prints ms per call, checks the proofs against LPK = 32)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fss_b200


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(3)

    def rnd(shape):
        return torch.randint(-2**31, 2**31 - 1, shape, dtype=torch.int64, device=dev, generator=g).to(torch.int32)

    for n, k in ((16, 8), (16, 256), (16, 2048), (12, 16384), (10, 65536), (20, 4)):
        ctx = fss_b200.Context("vdpf", n, "bytes")
        s0s, betas = rnd((k, 2, 4)), rnd((k, 4))
        s0s[..., 3] &= ~1
        betas[..., 3] &= ~1
        alphas = torch.randint(0, 1 << n, (k,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        cws, cs, ocws, _ = ctx.vdpf_gen(s0s, alphas, betas)
        seeds = s0s[:, 0].contiguous()
        row, want = [], None
        for lanes in ("32", "16", "8", "4", "2", "1", ""):
            if lanes:
                os.environ["FSSB200_VDPF_FINISH_LANES"] = lanes
            else:
                os.environ.pop("FSSB200_VDPF_FINISH_LANES", None)
            ya, pa = ctx.vdpf_eval_all(0, seeds, cws, cs, ocws)
            torch.cuda.synchronize()
            if want is None:
                want = (ya.clone(), pa.clone())
            assert torch.equal(ya, want[0]) and torch.equal(pa, want[1]), (n, k, lanes)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 2
            e0.record()
            for _ in range(reps):
                ctx.vdpf_eval_all(0, seeds, cws, cs, ocws)
            e1.record()
            torch.cuda.synchronize()
            row.append(f"{lanes or 'auto'}: {e0.elapsed_time(e1) / reps:.1f}")
            del ya, pa
        print(f"n={n} keys={k}  ms per EvalAll  " + "  ".join(row), flush=True)


if __name__ == "__main__":
    main()
