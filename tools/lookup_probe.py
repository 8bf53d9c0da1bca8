"""Lookup-path microbenchmarks (fssb200_microbench kinds 3, 7..12): does the texture pipe or the cached global-load
path add table-lookup bandwidth on top of the shared-memory pipe?  Prints one JSON object."""
import ctypes
import json
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fss_b200
from fss_b200 import _lib

NAMES = {3: "lds32_only(ilp8, legacy kernel)", 12: "lds32 x8", 7: "tex x4", 8: "lds32 x8 + tex x4", 9: "lds32 x8 + tex x2",
         10: "ldg.nc x4", 11: "lds32 x8 + ldg.nc x2"}


def main():
    assert torch.cuda.is_available()
    lib = _lib.lib
    out = {}
    for kind, name in NAMES.items():
        v = ctypes.c_double(0)
        rc = lib.fssb200_microbench(0, kind, ctypes.byref(v))
        out[name] = {"rc": rc, "lookups_per_s": v.value}
    lds = out["lds32 x8"]["lookups_per_s"]
    for name, row in out.items():
        row["vs_lds_only"] = row["lookups_per_s"] / lds if lds else None
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
