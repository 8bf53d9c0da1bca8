"""One small workload per invocation, for `ncu -k regex:<kernel> -s 2 -c 1` captures (profiles/).

  python tools/prof_one.py evalall_dpf|evalall_ht|evalall_dcf|gen_dcf|gen_dpf|grotto|c2|c3|ht|packed|vdpf|walk|relayout|lm
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fss_b200


def main():
    which = sys.argv[1]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(42)

    def rnd(shape):
        return torch.randint(-2**31, 2**31 - 1, shape, dtype=torch.int64, device=dev, generator=g).to(torch.int32)

    def inputs(k, n):
        s0s = rnd((k, 2, 4))
        s0s[..., 3] &= ~1
        betas = rnd((k, 4))
        betas[..., 3] &= ~1
        alphas = torch.randint(0, 1 << min(n, 62), (k,), dtype=torch.int64, device=dev, generator=g)
        return s0s, alphas, betas

    if which == "relayout":
        ctx = fss_b200.Context("dpf", 32, "bytes", prg="aes128_mmo")
        cws = rnd((1 << 22, 33, 8))
        for _ in range(4):
            ctx.relayout(cws)
    elif which in ("c2", "c3", "ht", "packed", "vdpf", "walk", "lm"):
        scheme, n, k, group = {"c2": ("dpf", 32, 1 << 22, "bytes"), "c3": ("dcf", 64, 1 << 21, "u128"),
                               "ht": ("halftree", 32, 1 << 20, "bytes"), "packed": ("dpf", 32, 1 << 21, "bytes"), "lm": ("dpf", 32, 1 << 22, "bytes"),
                               "vdpf": ("vdpf", 32, 1 << 20, "bytes"), "walk": ("grotto", 32, 1 << 20, "bytes")}[which]
        ctx = fss_b200.Context(scheme, n, group, prg="aes128_mmo")
        s0s, alphas, betas = inputs(k, n)
        xs = torch.randint(0, 1 << min(n, 62), (k,), dtype=torch.int64, device=dev, generator=g)
        if n <= 32:
            alphas, xs = alphas.to(torch.int32), xs.to(torch.int32)
        seeds0 = s0s[:, 0].contiguous()
        if scheme == "vdpf":
            cws, cs, ocws, _ = ctx.vdpf_gen(s0s, alphas, betas)
            for _ in range(3):
                ctx.vdpf_eval(0, seeds0, cws, cs, ocws, xs)
        elif scheme == "grotto":
            cws = ctx.gen(s0s, alphas)
            for _ in range(3):
                ctx.grotto_walk(0, seeds0, cws, xs)
        else:
            r = ctx.gen(s0s, alphas, betas)
            cws, ocws = r if scheme == "halftree" else (r, None)
            ys = torch.empty((k, 4), dtype=torch.int32, device=dev)
            if which == "lm":
                lay = ctx.relayout(cws)
                for _ in range(3):
                    ctx.eval_levelmajor(0, seeds0, lay, xs, ocws, out=ys)
            elif which == "packed":
                rows = ctx.pack_rows(cws.cpu()).to(dev)
                for _ in range(3):
                    ctx.eval_packed(0, seeds0, rows, xs, out=ys)
            else:
                for _ in range(3):
                    ctx.eval(0, seeds0, cws, xs, ocws, out=ys)
    elif which.startswith("evalall") or which == "grotto":
        scheme, n, k, group = {"evalall_dpf": ("dpf", 26, 2, "bytes"), "evalall_ht": ("halftree", 26, 2, "bytes"),
                               "evalall_dcf": ("dcf", 24, 2, "u128"), "grotto": ("grotto", 26, 2, "bytes")}[which]
        ctx = fss_b200.Context(scheme, n, group, prg="aes128_mmo")
        s0s, alphas, betas = inputs(k, n)
        r = ctx.gen(s0s, alphas) if scheme == "grotto" else ctx.gen(s0s, alphas, betas)
        cws, ocws = r if scheme == "halftree" else (r, None)
        seeds0 = s0s[:, 0].contiguous()
        if scheme == "grotto":
            for _ in range(3):
                ctx.grotto_expand(0, seeds0, cws)
        else:
            out = torch.empty((k, 1 << n, 4), dtype=torch.int32, device=dev)
            for _ in range(3):
                ctx.eval_all(0, seeds0, cws, ocws, out=out)
    else:
        scheme, n, k, group = {"gen_dcf": ("dcf", 64, 1 << 20, "u128"), "gen_dpf": ("dpf", 32, 1 << 21, "bytes")}[which]
        ctx = fss_b200.Context(scheme, n, group, prg="aes128_mmo")
        s0s, alphas, betas = inputs(k, n)
        for _ in range(3):
            ctx.gen(s0s, alphas, betas)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
