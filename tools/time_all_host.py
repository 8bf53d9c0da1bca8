#!/usr/bin/env python
"""fssb200_eval_all_host on small domains: whole keys per launch (default) against one key per launch
(`reserve_host(1)` = the loop this call ran before the last session).  Wall clock around the blocking call, CPU tensors
(pageable), best of 5; outputs compared."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import fss_b200
    g = torch.Generator().manual_seed(5)
    r = lambda s: torch.randint(-2 ** 31, 2 ** 31, s, dtype=torch.int64, generator=g).to(torch.int32)  # noqa: E731
    for scheme, n, k in (("dpf", 10, 4096), ("dpf", 14, 1024), ("halftree", 12, 2048), ("grotto", 12, 2048)):
        ctx = fss_b200.Context(scheme, n, "bytes" if scheme == "grotto" else "u64", prg="aes128_mmo")
        s0s, betas = r((k, 2, 4)), r((k, 4))
        s0s[:, :, 3] &= ~1
        betas[:, 3] &= ~1
        alphas = r((k,)) & ((1 << n) - 1)
        out = ctx.gen(s0s, alphas, None if scheme == "grotto" else betas)
        cws, ocws = out if scheme == "halftree" else (out, None)
        seeds = s0s[:, 0].contiguous()
        row = {"scheme": scheme, "in_bits": n, "keys": k}
        ys = {}
        for name, cap in (("whole_keys_per_launch", 0), ("one_key_per_launch", 1)):
            ctx.reserve_host(cap)
            best = None
            for _ in range(5):
                t0 = time.perf_counter()
                y = ctx.eval_all(0, seeds, cws, ocws)
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            ys[name] = y
            row[name + "_ms"] = round(best * 1e3, 3)
        row["identical"] = bool(torch.equal(ys["whole_keys_per_launch"], ys["one_key_per_launch"]))
        row["speedup"] = round(row["one_key_per_launch_ms"] / row["whole_keys_per_launch_ms"], 1)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
