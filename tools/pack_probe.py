"""Times fssb200_pack_rows alone (pageable vs pinned source, pageable vs pinned destination) and the host entry point
with and without packing.  Measurement tool."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fss_b200


def main():
    k, n = 1 << 20, 32
    ctx = fss_b200.Context("dpf", n, "bytes", prg="aes128_mmo")
    ctx.reserve_host(1 << 18, 0)
    out = {"threads": ctx.host_pack_threads(0)}
    src_pg = torch.randint(-2**31, 2**31 - 1, (k, n + 1, 8), dtype=torch.int64).to(torch.int32)
    src_pin = src_pg.pin_memory()
    rb = ctx.packed_row_bytes(0)
    import ctypes as C
    from fss_b200 import _lib as L
    h = ctx.handle(0)
    for sname, src in (("pageable", src_pg), ("pinned", src_pin)):
        for dname, dst in (("pageable", torch.empty((k, rb), dtype=torch.uint8)), ("pinned", torch.empty((k, rb), dtype=torch.uint8).pin_memory())):
            best = 1e9
            for _ in range(5):
                t0 = time.perf_counter()
                L.check(L.lib.fssb200_pack_rows(h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), k), "pack")
                best = min(best, time.perf_counter() - t0)
            out[f"pack_{sname}_to_{dname}_read_gbs"] = src.numel() * 4 / best / 1e9
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
