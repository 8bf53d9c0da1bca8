"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends of the two CPU checkers.

* ``Orc``  : oracle/liboracle.so, the plain-C restatement (oracle/fss_oracle.c).
* ``Ref``  : oracle/_ref/libfssref.so, the UNMODIFIED reference headers compiled by
             oracle/Makefile (present only after ``make -C oracle ref`` in a container that
             has /root/reference; the built .so travels to the GPU box).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this package; the product (fss_b200/) never does.

Both classes expose the same numpy-in / numpy-out methods:
  gen(s0s[K,2,4]u32, alphas[K] int, betas[K,4]u32) -> cws[K,ncw,8]u32 (, ocws[K,4]u32)
  eval(party, seeds[K,4], cws, xs[K] int, ocws=None) -> ys[K,4]u32
  evalall(party, seeds, cws, ocws=None, leaf_begin=0, leaf_count=0) -> ys[K,L,4]u32 | [K,L]u8
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SCHEME = {"dpf": 0, "dcf": 1, "halftree": 2, "grotto": 3, "vdpf": 4}
GROUP = {"bytes": 0, "u8": 1, "u16": 2, "u32": 3, "u64": 4, "u128": 5}
PRG = {"aes128_mmo": 0, "chacha": 1, "aes128_mmo_raw": 2}
PRED = {"lt": 0, "gt": 1}
HASH = {"blake3": 0, "sha256": 1}

# fixtures used by the reference's own tests / samples
AES_KEYS = bytes(range(1, 17)) + bytes(range(16, 0, -1)) + bytes(
    [1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8]) + bytes([8, 8, 7, 7, 6, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1])
CHACHA_NONCE = np.array([0x12345678, 0x9ABCDEF0], dtype=np.uint32).tobytes()
HASH_KEY_SAMPLE = np.array([0x12345678, 0x9ABCDEF0, 0x13572468, 0x2468ACE0], dtype=np.uint32).tobytes()
HASH_KEY_BENCH = np.array([0x12345678, 0x9ABCDEF0, 0x0FEDCBA9, 0x87654321], dtype=np.uint32).tobytes()
# VDPF: IVs of the XorHash / Hash Blake3 plugins -- the constants of the reference's tests
# (src/vdpf_test.cu:33-42, samples/vdpf_cpu.cu) for XorHash, a second pattern for Hash
HASH_IVS = (np.array([0x11111111, 0x22222222, 0x33333333, 0x44444444, 0x55555555, 0x66666666, 0x77777777, 0x88888888],
                     dtype=np.uint32).tobytes()
            + np.array([0x99999999, 0xAAAAAAAA, 0xBBBBBBBB, 0xCCCCCCCC, 0xDDDDDDDD, 0xEEEEEEEE, 0xFFFFFFFF, 0x01234567],
                       dtype=np.uint32).tobytes())


class CParams(C.Structure):
    """Mirror of ``fssb200_params`` (include/fssb200.h)."""
    _fields_ = [("scheme", C.c_int32), ("in_bits", C.c_int32), ("in_bytes", C.c_int32), ("group", C.c_int32),
                ("mod_lo", C.c_uint64), ("mod_hi", C.c_uint64), ("prg", C.c_int32), ("pred", C.c_int32),
                ("prg_key", C.c_uint8 * 64), ("hash_key", C.c_uint8 * 16), ("device", C.c_int32),
                ("hash", C.c_int32), ("hash_iv", C.c_uint8 * 64)]


class CRefParams(C.Structure):
    _fields_ = [("in_bytes", C.c_int32), ("prg_key", C.c_uint8 * 64), ("hash_key", C.c_uint8 * 16)]


class CRefSel(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("in_bits", C.c_int32), ("group", C.c_int32), ("prg", C.c_int32),
                ("pred", C.c_int32), ("pad", C.c_int32), ("mod_lo", C.c_uint64), ("mod_hi", C.c_uint64)]


def in_bytes_for(in_bits: int) -> int:
    """fss_crypto/_jit.py:57-62: uint32_t / uint64_t / __uint128_t by in_bits."""
    return 4 if in_bits <= 32 else (8 if in_bits <= 64 else 16)


@dataclass
class Params:
    scheme: str = "dpf"
    in_bits: int = 32
    group: str = "bytes"
    mod: int = 0
    prg: str = "aes128_mmo"
    pred: str = "lt"
    prg_key: bytes = b""
    hash_key: bytes = HASH_KEY_SAMPLE
    in_bytes: int = 0
    hash_iv: bytes = HASH_IVS
    hash: tuple = ("blake3", "blake3")   # VDPF (XorHash, Hash) plugins: "blake3" | "sha256"

    def __post_init__(self):
        if not self.in_bytes:
            self.in_bytes = in_bytes_for(self.in_bits)
        if not self.prg_key:
            self.prg_key = AES_KEYS if self.prg.startswith("aes") else CHACHA_NONCE
        if self.group == "u128" and self.mod == 0:
            self.mod = 1 << 127

    @property
    def ncw(self) -> int:
        return self.in_bits if self.scheme in ("halftree", "vdpf") else self.in_bits + 1

    @property
    def mul(self) -> int:
        return {"dpf": 2, "dcf": 4, "halftree": 1, "grotto": 2, "vdpf": 2}[self.scheme]

    def c(self) -> CParams:
        p = CParams()
        p.scheme, p.in_bits, p.in_bytes, p.group = SCHEME[self.scheme], self.in_bits, self.in_bytes, GROUP[self.group]
        p.mod_lo, p.mod_hi = self.mod & (2 ** 64 - 1), self.mod >> 64
        p.prg, p.pred = PRG[self.prg], PRED[self.pred]
        key = self.prg_key.ljust(64, b"\0")
        for i in range(64):
            p.prg_key[i] = key[i]
        for i in range(16):
            p.hash_key[i] = self.hash_key[i]
        C.memmove(p.hash_iv, bytes(self.hash_iv), 64)
        hx, hh = (self.hash, self.hash) if isinstance(self.hash, str) else self.hash
        p.hash = HASH[hx] | (HASH[hh] << 8)
        return p


def pack_ints(vals, in_bytes: int) -> np.ndarray:
    """Python ints / array -> little-endian ``In[K]`` as a uint8 [K, in_bytes] array."""
    if isinstance(vals, np.ndarray) and vals.dtype.kind == "u" and vals.dtype.itemsize == in_bytes:
        return np.ascontiguousarray(vals).view(np.uint8).reshape(len(vals), in_bytes)
    out = np.zeros((len(vals), in_bytes), dtype=np.uint8)
    for i, v in enumerate(vals):
        out[i] = np.frombuffer(int(v).to_bytes(in_bytes, "little"), dtype=np.uint8)
    return out


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a if shape is None else a.reshape(shape)


class _Base:
    def leaf_bytes(self, p: Params) -> int:
        return 1 if p.scheme == "grotto" else 16


class Orc(_Base):
    """The plain-C restatement."""
    kind = "port"

    def __init__(self, build: bool = True):
        path = os.path.join(_HERE, "liboracle.so")
        if build and (not os.path.exists(path)
                      or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "fss_oracle.c"))):
            subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
        self.lib = C.CDLL(path)

    def prg_gen(self, p: Params, mul: int, seeds: np.ndarray) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        out = np.zeros((len(seeds), mul, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_prg_gen(C.byref(cp), mul, C.c_size_t(len(seeds)), _vp(seeds), _vp(out))
        assert rc == 0, rc
        return out

    def gen(self, p: Params, s0s, alphas, betas=None, threads: int = 1):
        s0s = _u32(s0s, (-1, 2, 4))
        k = len(s0s)
        al = pack_ints(alphas, p.in_bytes)
        be = None if betas is None else _u32(betas, (k, 4))
        cws = np.zeros((k, p.ncw, 8), dtype=np.uint32)
        ocws = np.zeros((k, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_gen(C.byref(cp), C.c_size_t(k), _vp(s0s), _vp(al), _vp(be), _vp(cws), _vp(ocws), threads)
        assert rc == 0, rc
        return (cws, ocws) if p.scheme == "halftree" else cws

    def eval(self, p: Params, party: int, seeds, cws, xs, ocws=None, threads: int = 1) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws = _u32(cws, (k, p.ncw, 8))
        xb = pack_ints(xs, p.in_bytes)
        oc = None if ocws is None else _u32(ocws, (k, 4))
        ys = np.zeros((k, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_eval(C.byref(cp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(oc), _vp(xb), _vp(ys),
                               threads)
        assert rc == 0, rc
        return ys

    def evalall(self, p: Params, party: int, seeds, cws, ocws=None, leaf_begin: int = 0, leaf_count: int = 0,
                threads: int = 1) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws = _u32(cws, (k, p.ncw, 8))
        oc = None if ocws is None else _u32(ocws, (k, 4))
        cnt = leaf_count or ((1 << p.in_bits) - leaf_begin)
        ys = np.zeros((k, cnt), dtype=np.uint8) if p.scheme == "grotto" else np.zeros((k, cnt, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_evalall(C.byref(cp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(oc), _vp(ys),
                                  C.c_uint64(leaf_begin), C.c_uint64(leaf_count), threads)
        assert rc == 0, rc
        return ys

    def grotto_expand(self, p: Params, party: int, seeds, cws, leaf_begin: int = 0, leaf_count: int = 0,
                      threads: int = 1) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws = _u32(cws, (k, p.ncw, 8))
        cnt = leaf_count or ((1 << p.in_bits) - leaf_begin)
        t = np.zeros((k, cnt), dtype=np.uint8)
        cp = p.c()
        rc = self.lib.orc_grotto_expand(C.byref(cp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(t),
                                        C.c_uint64(leaf_begin), C.c_uint64(leaf_count), threads)
        assert rc == 0, rc
        return t

    def grotto_preprocess(self, p: Params, party: int, seeds, cws, threads: int = 1) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws = _u32(cws, (k, p.ncw, 8))
        pt = np.zeros((k, (2 << p.in_bits) - 1), dtype=np.uint8)
        cp = p.c()
        rc = self.lib.orc_grotto_preprocess(C.byref(cp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(pt), threads)
        assert rc == 0, rc
        return pt

    def grotto_lookup(self, p: Params, pt: np.ndarray, xs) -> np.ndarray:
        pt = np.ascontiguousarray(pt, dtype=np.uint8)
        k = len(pt)
        xb = pack_ints(xs, p.in_bytes)
        ys = np.zeros(k, dtype=np.uint8)
        cp = p.c()
        rc = self.lib.orc_grotto_lookup(C.byref(cp), C.c_size_t(k), _vp(pt), _vp(xb), _vp(ys))
        assert rc == 0, rc
        return ys

    def relayout(self, p: Params, cws):
        cws = _u32(cws, (-1, p.ncw, 8))
        k, n = len(cws), p.in_bits
        cw_s = np.zeros((n, k, 4), dtype=np.uint32)
        cw_v = np.zeros((n, k, 4), dtype=np.uint32)
        extra = np.zeros(((n + 31) // 32, k), dtype=np.uint32)
        out_cw = np.zeros((k, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_relayout(C.byref(cp), C.c_size_t(k), _vp(cws), _vp(cw_s), _vp(cw_v), _vp(extra),
                                   _vp(out_cw))
        assert rc == 0, rc
        return cw_s, cw_v, extra, out_cw

    # ---- VDPF (vdpf.cuh) ---------------------------------------------------------------------------------
    def hash(self, p: Params, which: int, msgs) -> np.ndarray:
        """which 0: XorHash over (a, b) pairs [K,2,4] -> [K,4,4]; 1: Hash over [K,4,4] -> [K,2,4]."""
        msgs = _u32(msgs, (-1, 2, 4) if which == 0 else (-1, 4, 4))
        out = np.zeros((len(msgs), 4 if which == 0 else 2, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_hash(C.byref(cp), which, C.c_size_t(len(msgs)), _vp(msgs), _vp(out))
        assert rc == 0, rc
        return out

    def vdpf_gen(self, p: Params, s0s, alphas, betas, threads: int = 1):
        """-> cws[K,n,8], cs[K,4,4], ocws[K,4], status[K] (1 = resample the seeds; ocw not written)."""
        s0s = _u32(s0s, (-1, 2, 4))
        k = len(s0s)
        al, be = pack_ints(alphas, p.in_bytes), _u32(betas, (k, 4))
        cws, cs = np.zeros((k, p.ncw, 8), dtype=np.uint32), np.zeros((k, 4, 4), dtype=np.uint32)
        ocws, status = np.zeros((k, 4), dtype=np.uint32), np.zeros(k, dtype=np.int32)
        cp = p.c()
        rc = self.lib.orc_vdpf_gen(C.byref(cp), C.c_size_t(k), _vp(s0s), _vp(al), _vp(be), _vp(cws), _vp(cs),
                                   _vp(ocws), _vp(status), threads)
        assert rc == 0, rc
        return cws, cs, ocws, status

    def vdpf_eval(self, p: Params, party: int, seeds, cws, cs, ocws, xs, threads: int = 1):
        """-> ys[K,4], pi_tildes[K,4,4]."""
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws, cs, ocws = _u32(cws, (k, p.ncw, 8)), _u32(cs, (k, 4, 4)), _u32(ocws, (k, 4))
        xb = pack_ints(xs, p.in_bytes)
        ys, pis = np.zeros((k, 4), dtype=np.uint32), np.zeros((k, 4, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_vdpf_eval(C.byref(cp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(cs), _vp(ocws),
                                    _vp(xb), _vp(ys), _vp(pis), threads)
        assert rc == 0, rc
        return ys, pis

    def vdpf_prove(self, p: Params, pi_tildes, cs) -> np.ndarray:
        """pi_tildes[K,m,4,4], cs[K,4,4] -> pi[K,4,4]."""
        cs = _u32(cs, (-1, 4, 4))
        k = len(cs)
        pts = _u32(pi_tildes).reshape(k, -1, 4, 4)
        pis = np.zeros((k, 4, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_vdpf_prove(C.byref(cp), C.c_size_t(k), C.c_size_t(pts.shape[1]), _vp(pts), _vp(cs), _vp(pis))
        assert rc == 0, rc
        return pis

    def vdpf_evalall(self, p: Params, party: int, seeds, cws, cs, ocws, threads: int = 1):
        """-> ys[K,2^n,4], pi[K,4,4]."""
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws, cs, ocws = _u32(cws, (k, p.ncw, 8)), _u32(cs, (k, 4, 4)), _u32(ocws, (k, 4))
        ys, pis = np.zeros((k, 1 << p.in_bits, 4), dtype=np.uint32), np.zeros((k, 4, 4), dtype=np.uint32)
        cp = p.c()
        rc = self.lib.orc_vdpf_evalall(C.byref(cp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(cs), _vp(ocws),
                                       _vp(ys), _vp(pis), threads)
        assert rc == 0, rc
        return ys, pis

    def group_add(self, p: Params, a, b) -> np.ndarray:
        a, b = _u32(a, (-1, 4)), _u32(b, (-1, 4))
        out = np.zeros_like(a)
        cp = p.c()
        rc = self.lib.orc_group_add(C.byref(cp), C.c_size_t(len(a)), _vp(a), _vp(b), _vp(out))
        assert rc == 0, rc
        return out


class Ref(_Base):
    """The compiled reference (oracle/_ref/libfssref.so)."""
    kind = "reference"

    @staticmethod
    def path() -> str:
        return os.path.join(_HERE, "_ref", "libfssref.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.path())

    def __init__(self):
        self.lib = C.CDLL(self.path())

    def host_threads(self) -> int:
        return int(self.lib.ref_host_threads())

    @staticmethod
    def _sel(p: Params) -> CRefSel:
        hx, hh = (p.hash, p.hash) if isinstance(p.hash, str) else p.hash
        # (the reference shim instantiates XorHash == Hash only; a mixed pair selects nothing)
        return CRefSel(SCHEME[p.scheme], p.in_bits, GROUP[p.group], PRG[p.prg], PRED[p.pred],
                       HASH[hx] if hx == hh else 99, p.mod & (2 ** 64 - 1), p.mod >> 64)

    @staticmethod
    def _rp(p: Params) -> CRefParams:
        r = CRefParams()
        r.in_bytes = p.in_bytes
        key = p.prg_key.ljust(64, b"\0")
        for i in range(64):
            r.prg_key[i] = key[i]
        for i in range(16):
            r.hash_key[i] = p.hash_key[i]
        return r

    def supported(self, p: Params) -> bool:
        s = self._sel(p)
        return bool(self.lib.ref_supported(C.byref(s)))

    def prg_gen(self, p: Params, mul: int, seeds: np.ndarray, prg_tag: int | None = None) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        out = np.zeros((len(seeds), mul, 4), dtype=np.uint32)
        rp = self._rp(p)
        tag = PRG[p.prg] if prg_tag is None else prg_tag
        rc = self.lib.ref_prg_gen(C.byref(rp), tag, mul, C.c_size_t(len(seeds)), _vp(seeds), _vp(out))
        assert rc == 0, rc
        return out

    def gen(self, p: Params, s0s, alphas, betas=None, threads: int = 1):
        s0s = _u32(s0s, (-1, 2, 4))
        k = len(s0s)
        al = pack_ints(alphas, p.in_bytes)
        be = None if betas is None else _u32(betas, (k, 4))
        cws = np.zeros((k, p.ncw, 8), dtype=np.uint32)
        ocws = np.zeros((k, 4), dtype=np.uint32)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_gen(C.byref(s), C.byref(rp), C.c_size_t(k), _vp(s0s), _vp(al), _vp(be), _vp(cws),
                              _vp(ocws), threads)
        assert rc == 0, f"reference instantiation missing for {p}"
        return (cws, ocws) if p.scheme == "halftree" else cws

    def eval(self, p: Params, party: int, seeds, cws, xs, ocws=None, threads: int = 1) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws = _u32(cws, (k, p.ncw, 8))
        xb = pack_ints(xs, p.in_bytes)
        oc = None if ocws is None else _u32(ocws, (k, 4))
        ys = np.zeros((k, 4), dtype=np.uint32)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_eval(C.byref(s), C.byref(rp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(oc), _vp(xb),
                               _vp(ys), threads)
        assert rc == 0, f"reference instantiation missing for {p}"
        return ys

    def evalall(self, p: Params, party: int, seeds, cws, ocws=None, leaf_begin: int = 0, leaf_count: int = 0,
                threads: int = 1) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws = _u32(cws, (k, p.ncw, 8))
        oc = None if ocws is None else _u32(ocws, (k, 4))
        n_leaves = 1 << p.in_bits
        ys = np.zeros((k, n_leaves), dtype=np.uint8) if p.scheme == "grotto" else np.zeros((k, n_leaves, 4),
                                                                                          dtype=np.uint32)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_evalall(C.byref(s), C.byref(rp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(oc),
                                  _vp(ys), threads)
        assert rc == 0, f"reference instantiation missing for {p}"
        cnt = leaf_count or (n_leaves - leaf_begin)
        if leaf_begin or cnt != n_leaves:
            assert p.scheme != "grotto" or leaf_begin == 0
            ys = np.ascontiguousarray(ys[:, leaf_begin:leaf_begin + cnt])
        return ys

    def grotto_preprocess(self, p: Params, party: int, seeds, cws, threads: int = 1) -> np.ndarray:
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws = _u32(cws, (k, p.ncw, 8))
        pt = np.zeros((k, (2 << p.in_bits) - 1), dtype=np.uint8)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_grotto_preprocess(C.byref(s), C.byref(rp), party, C.c_size_t(k), _vp(seeds), _vp(cws),
                                            _vp(pt), threads)
        assert rc == 0, f"reference instantiation missing for {p}"
        return pt

    def grotto_lookup(self, p: Params, pt: np.ndarray, xs) -> np.ndarray:
        pt = np.ascontiguousarray(pt, dtype=np.uint8)
        k = len(pt)
        xb = pack_ints(xs, p.in_bytes)
        ys = np.zeros(k, dtype=np.uint8)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_grotto_lookup(C.byref(s), C.byref(rp), C.c_size_t(k), _vp(pt), _vp(xb), _vp(ys))
        assert rc == 0, f"reference instantiation missing for {p}"
        return ys


    # ---- VDPF (oracle/ref_vdpf.cpp) -------------------------------------------------------------------------
    @staticmethod
    def _ivs(p: Params):
        return (C.c_uint8 * 64).from_buffer_copy(bytes(p.hash_iv))

    def vdpf_supported(self, p: Params) -> bool:
        s = self._sel(p)
        return bool(self.lib.ref_vdpf_supported(C.byref(s)))

    def hash(self, p: Params, which: int, msgs) -> np.ndarray:
        msgs = _u32(msgs, (-1, 2, 4) if which == 0 else (-1, 4, 4))
        out = np.zeros((len(msgs), 4 if which == 0 else 2, 4), dtype=np.uint32)
        iv = (C.c_uint8 * 32).from_buffer_copy(bytes(p.hash_iv)[32 * which:32 * which + 32])
        hname = (p.hash if isinstance(p.hash, str) else p.hash[which])
        fn = self.lib.ref_sha256 if hname == "sha256" else self.lib.ref_blake3
        rc = fn(iv, which, C.c_size_t(len(msgs)), _vp(msgs), _vp(out))
        assert rc == 0, rc
        return out

    def vdpf_gen(self, p: Params, s0s, alphas, betas, threads: int = 1):
        s0s = _u32(s0s, (-1, 2, 4))
        k = len(s0s)
        al, be = pack_ints(alphas, p.in_bytes), _u32(betas, (k, 4))
        cws, cs = np.zeros((k, p.ncw, 8), dtype=np.uint32), np.zeros((k, 4, 4), dtype=np.uint32)
        ocws, status = np.zeros((k, 4), dtype=np.uint32), np.zeros(k, dtype=np.int32)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_vdpf_gen(C.byref(s), C.byref(rp), self._ivs(p), C.c_size_t(k), _vp(s0s), _vp(al), _vp(be),
                                   _vp(cws), _vp(cs), _vp(ocws), _vp(status), threads)
        assert rc == 0, f"reference VDPF instantiation missing for {p}"
        return cws, cs, ocws, status

    def vdpf_eval(self, p: Params, party: int, seeds, cws, cs, ocws, xs, threads: int = 1):
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws, cs, ocws = _u32(cws, (k, p.ncw, 8)), _u32(cs, (k, 4, 4)), _u32(ocws, (k, 4))
        xb = pack_ints(xs, p.in_bytes)
        ys, pis = np.zeros((k, 4), dtype=np.uint32), np.zeros((k, 4, 4), dtype=np.uint32)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_vdpf_eval(C.byref(s), C.byref(rp), self._ivs(p), party, C.c_size_t(k), _vp(seeds), _vp(cws),
                                    _vp(cs), _vp(ocws), _vp(xb), _vp(ys), _vp(pis), threads)
        assert rc == 0, f"reference VDPF instantiation missing for {p}"
        return ys, pis

    def vdpf_prove(self, p: Params, pi_tildes, cs) -> np.ndarray:
        cs = _u32(cs, (-1, 4, 4))
        k = len(cs)
        pts = _u32(pi_tildes).reshape(k, -1, 4, 4)
        pis = np.zeros((k, 4, 4), dtype=np.uint32)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_vdpf_prove(C.byref(s), C.byref(rp), self._ivs(p), C.c_size_t(k), C.c_size_t(pts.shape[1]),
                                     _vp(pts), _vp(cs), _vp(pis))
        assert rc == 0, f"reference VDPF instantiation missing for {p}"
        return pis

    def vdpf_evalall(self, p: Params, party: int, seeds, cws, cs, ocws, threads: int = 1):
        seeds = _u32(seeds, (-1, 4))
        k = len(seeds)
        cws, cs, ocws = _u32(cws, (k, p.ncw, 8)), _u32(cs, (k, 4, 4)), _u32(ocws, (k, 4))
        ys, pis = np.zeros((k, 1 << p.in_bits, 4), dtype=np.uint32), np.zeros((k, 4, 4), dtype=np.uint32)
        s, rp = self._sel(p), self._rp(p)
        rc = self.lib.ref_vdpf_evalall(C.byref(s), C.byref(rp), self._ivs(p), party, C.c_size_t(k), _vp(seeds),
                                       _vp(cws), _vp(cs), _vp(ocws), _vp(ys), _vp(pis), threads)
        assert rc == 0, f"reference VDPF instantiation missing for {p}"
        return ys, pis


def synth_inputs(p: Params, nkeys: int, seed: int = 42, alpha_hit_every: int = 16):
    """Seeded synthetic keys/inputs in the spirit of src/bench_gpu.cu:254-262: random clamped seeds,
    uniform alpha / x / beta, every ``alpha_hit_every``-th x forced to alpha (beta branch exercised)."""
    rng = np.random.default_rng(seed)
    s0s = rng.integers(0, 2 ** 32, size=(nkeys, 2, 4), dtype=np.uint64).astype(np.uint32)
    s0s[:, :, 3] &= 0xFFFFFFFE
    betas = rng.integers(0, 2 ** 32, size=(nkeys, 4), dtype=np.uint64).astype(np.uint32)
    betas[:, 3] &= 0xFFFFFFFE
    nb = p.in_bytes
    raw = rng.integers(0, 256, size=(2, nkeys, nb), dtype=np.uint64).astype(np.uint8)
    mask = (1 << p.in_bits) - 1
    def to_ints(a):
        return [int.from_bytes(a[i].tobytes(), "little") & mask for i in range(nkeys)]
    alphas, xs = to_ints(raw[0]), to_ints(raw[1])
    for i in range(0, nkeys, alpha_hit_every):
        xs[i] = alphas[i]
    return s0s, alphas, betas, xs
