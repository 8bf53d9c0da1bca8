"""Loader of tests/golden/golden_v1.{json,npz} (generated from the reference by oracle/make_golden.py)."""
import json
import os

import numpy as np

from oracle import Params

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case:
    def __init__(self, meta, arrays):
        self.meta, self.name = meta, meta["name"]
        self.p = Params(scheme=meta["scheme"], in_bits=meta["in_bits"], group=meta["group"], mod=int(meta["mod"]),
                        prg=meta["prg"], pred=meta["pred"], prg_key=bytes.fromhex(meta["prg_key"]),
                        hash_key=bytes.fromhex(meta["hash_key"]), in_bytes=meta["in_bytes"])
        self.alphas = [int(a) for a in meta["alphas"]]
        self.xs = [int(x) for x in meta["xs"]]
        self._arrays = arrays

    def __getitem__(self, key):
        return self._arrays[f"{self.name}/{key}"]

    def has(self, key):
        return f"{self.name}/{key}" in self._arrays

    @property
    def betas(self):
        return None if self.p.scheme == "grotto" else self["betas"]

    def masked_cws(self, cws):
        """Correction words with the bytes the reference leaves unspecified zeroed (struct padding of
        Dpf::Cw / HalfTreeDpf::Cw, and the whole second half of the aggregate-assigned entry [n],
        dpf.cuh:158)."""
        k = cws.shape[0]
        b = np.ascontiguousarray(cws).copy().view(np.uint8).reshape(k, self.p.ncw, 32)
        if self.p.scheme != "dcf":
            b[:, :, 17:] = 0
            if self.p.scheme != "halftree":
                b[:, -1, 16:] = 0
        return b


class Golden:
    def __init__(self):
        with open(os.path.join(_DIR, "golden_v1.json")) as f:
            self.manifest = json.load(f)
        self.arrays = np.load(os.path.join(_DIR, "golden_v1.npz"))
        self.cases = [Case(m, self.arrays) for m in self.manifest["cases"]]

    def by_name(self, name):
        for c in self.cases:
            if c.name == name:
                return c
        raise KeyError(name)
