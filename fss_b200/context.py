"""Torch-facing wrapper of the C ABI (include/fssb200.h): one ``Context`` per parameter set.

PyTorch is plumbing here (device memory, streams); every method below ends in a call into
libfssb200.so -- there is no Python / CPU evaluation path.

Tensor conventions (the reference binding's, fss_crypto/_csrc/dpf_binding_impl.cuh:44-46):
  seeds  (N, 4) int32       s0s (N, 2, 4) int32       betas (N, 4) int32
  cws    (N, ncw, 8) int32  row = 8 int32: words 0-3 ``s``; word 4 = ``tr`` (DPF) / words 4-7 ``v`` (DCF)
  ys     (N, 4) int32       eval_all: (N, L, 4) int32, Grotto (N, L) uint8
  xs / alphas: Python ints, or an int32 / int64 / uint8[N, in_bytes] tensor (little-endian ``In``)
CUDA tensors run stream-ordered on torch's current stream; CPU tensors go through the
``*_host`` entry points (chunked H2D / kernel / D2H pipeline inside the library).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import torch

from . import _lib as L

_SCHEMES = {"dpf": L.SCHEME_DPF, "dcf": L.SCHEME_DCF, "halftree": L.SCHEME_HALFTREE, "grotto": L.SCHEME_GROTTO,
            "vdpf": L.SCHEME_VDPF}
_GROUPS = {"bytes": L.GROUP_BYTES, "u8": L.GROUP_U8, "u16": L.GROUP_U16, "u32": L.GROUP_U32, "u64": L.GROUP_U64,
           "u128": L.GROUP_U128}
_PRGS = {"aes128_mmo": L.PRG_AES128_MMO, "chacha": L.PRG_CHACHA}
_PREDS = {"lt": L.PRED_LT, "gt": L.PRED_GT}

# fixtures of the reference's samples/tests (samples/dpf_dcf_cpu.cu:39-42,98-101; src/dpf_test.cu:35)
DEFAULT_AES_KEYS = bytes(range(1, 17)) + bytes(range(16, 0, -1)) + bytes(
    [1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8]) + bytes([8, 8, 7, 7, 6, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1])
DEFAULT_CHACHA_NONCE = (0x12345678).to_bytes(4, "little") + (0x9ABCDEF0).to_bytes(4, "little")
DEFAULT_HASH_KEY = b"".join(v.to_bytes(4, "little") for v in (0x12345678, 0x9ABCDEF0, 0x0FEDCBA9, 0x87654321))
# VDPF: IVs of the XorHash / Hash Blake3 plugins (first one = the reference tests' constant, src/vdpf_test.cu:35-36)
_HASHES = {"blake3": 0, "sha256": 1}
DEFAULT_HASH_IVS = b"".join(v.to_bytes(4, "little") for v in (
    0x11111111, 0x22222222, 0x33333333, 0x44444444, 0x55555555, 0x66666666, 0x77777777, 0x88888888,
    0x99999999, 0xAAAAAAAA, 0xBBBBBBBB, 0xCCCCCCCC, 0xDDDDDDDD, 0xEEEEEEEE, 0xFFFFFFFF, 0x01234567))

IntLike = Union[int, Sequence[int], torch.Tensor]


def in_bytes_for(in_bits: int) -> int:
    """``In`` by in_bits as the reference binding picks it (fss_crypto/_jit.py:57-62)."""
    return 4 if in_bits <= 32 else (8 if in_bits <= 64 else 16)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Context:
    def __init__(self, scheme: str, in_bits: int, group: str = "bytes", mod: int = 0, prg: str = "aes128_mmo",
                 pred: str = "lt", prg_key: Optional[bytes] = None, hash_key: Optional[bytes] = None,
                 in_bytes: Optional[int] = None, hash_iv: Optional[bytes] = None, hash: "str | tuple[str, str]" = "blake3"):
        self.scheme, self.in_bits, self.group, self.prg, self.pred = scheme, in_bits, group, prg, pred
        if group == "u128" and mod == 0:
            mod = 1 << 127
        self.mod = mod
        self.in_bytes = in_bytes or in_bytes_for(in_bits)
        self.prg_key = prg_key if prg_key is not None else (
            DEFAULT_AES_KEYS if prg == "aes128_mmo" else DEFAULT_CHACHA_NONCE)
        self.hash_key = hash_key if hash_key is not None else DEFAULT_HASH_KEY
        self.hash_iv = hash_iv if hash_iv is not None else DEFAULT_HASH_IVS
        if len(self.hash_iv) != 64:
            raise ValueError("hash_iv must be 64 bytes (XorHash IV || Hash IV)")
        # VDPF hash plugins (XorHash, Hash): "blake3" (hash/blake3.cuh) or "sha256" (hash/sha256.cuh; its 16-byte key is
        # the first 16 bytes of the plugin's 32-byte hash_iv slot); one name = both plugins
        hx, hh = (hash, hash) if isinstance(hash, str) else hash
        if hx not in _HASHES or hh not in _HASHES:
            raise ValueError(f"hash must be one of {sorted(_HASHES)}")
        self.hash_plugins = (hx, hh)
        self.ncw = in_bits if scheme in ("halftree", "vdpf") else in_bits + 1
        self.mul = {"dpf": 2, "dcf": 4, "halftree": 1, "grotto": 2, "vdpf": 2}[scheme]
        self._handles: dict[int, C.c_void_p] = {}

    # ---- context handles -------------------------------------------------------------------------
    def _params(self, device: int) -> L.Params:
        p = L.Params()
        p.scheme, p.in_bits, p.in_bytes, p.group = _SCHEMES[self.scheme], self.in_bits, self.in_bytes, _GROUPS[
            self.group]
        p.mod_lo, p.mod_hi = self.mod & (2 ** 64 - 1), self.mod >> 64
        p.prg, p.pred, p.device = _PRGS[self.prg], _PREDS[self.pred], device
        key = bytes(self.prg_key).ljust(64, b"\0")
        C.memmove(p.prg_key, key, 64)
        C.memmove(p.hash_key, bytes(self.hash_key), 16)
        C.memmove(p.hash_iv, bytes(self.hash_iv), 64)
        p.hash = _HASHES[self.hash_plugins[0]] | (_HASHES[self.hash_plugins[1]] << 8)
        return p

    def handle(self, device: Optional[int] = None) -> C.c_void_p:
        if device is None:
            device = torch.cuda.current_device()
        h = self._handles.get(device)
        if h is None:
            p = self._params(device)
            h = C.c_void_p()
            L.check(L.lib.fssb200_ctx_create(C.byref(p), C.byref(h)), "fssb200_ctx_create")
            self._handles[device] = h
        return h

    def close(self) -> None:
        for h in self._handles.values():
            L.lib.fssb200_ctx_destroy(h)
        self._handles.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self, device: Optional[int] = None) -> int:
        return int(L.lib.fssb200_ctx_launch_count(self.handle(device)))

    def granule(self, device: Optional[int] = None) -> int:
        return int(L.lib.fssb200_eval_all_granule(self.handle(device)))

    def reserve_host(self, max_keys_per_chunk: int = 0, device: Optional[int] = None) -> None:
        """Preferred keys per chunk of the host-buffer calls (0 = library default).  Allocates nothing: the
        staging arenas belong to a process-wide pool and are sized per call (include/fssb200.h)."""
        L.check(L.lib.fssb200_ctx_reserve_host(self.handle(device), max_keys_per_chunk), "fssb200_ctx_reserve_host")

    def set_host_mode(self, mode: int, device: Optional[int] = None) -> None:
        """eval() on CPU tensors: 0 = adaptive pack / direct pipeline, 1 = reference layout only, 2 = staged only."""
        L.check(L.lib.fssb200_ctx_set_host_mode(self.handle(device), mode), "fssb200_ctx_set_host_mode")

    def host_stats(self, device: Optional[int] = None) -> dict:
        """Of the last host-buffer eval(): keys that crossed the link packed / in the reference layout, threads."""
        p, d, t = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        L.check(L.lib.fssb200_ctx_host_stats(self.handle(device), C.byref(p), C.byref(d), C.byref(t)),
                "fssb200_ctx_host_stats")
        return {"packed_keys": int(p.value), "direct_keys": int(d.value), "threads": int(t.value)}

    def _ensure_host(self, device: int) -> None:  # (0.1 needed an explicit arena; calls size their own now)
        pass

    # ---- helpers -----------------------------------------------------------------------------------------
    def in_tensor(self, vals: IntLike, device: torch.device) -> torch.Tensor:
        """Domain values -> contiguous little-endian ``In[N]`` tensor on ``device``."""
        nb = self.in_bytes
        if isinstance(vals, torch.Tensor):
            t = vals
            if t.dtype == torch.uint8 and t.dim() == 2 and t.shape[1] == nb:
                return t.contiguous().to(device)
            if t.dim() != 1 or t.dtype not in (torch.int32, torch.int64):
                raise TypeError(f"domain values must be int32/int64 (N,) or uint8 (N, {nb}), got "
                                f"shape {tuple(t.shape)} dtype {t.dtype}")
            if nb == 4:
                return t.to(torch.int32).contiguous().to(device)
            if nb == 8:
                return t.to(torch.int64).contiguous().to(device)
            if nb == 16:
                t64 = t.to(torch.int64)
                return torch.stack([t64, torch.zeros_like(t64)], dim=1).contiguous().to(device)
            return t.to(torch.int64).view(torch.uint8).reshape(-1, 8)[:, :nb].contiguous().to(device)
        if isinstance(vals, int):
            vals = [vals]
        top = 1 << self.in_bits
        for v in vals:  # (tensors are not range-checked: that would cost a device synchronisation per call)
            if not 0 <= int(v) < top:
                raise ValueError(f"domain value {int(v)} outside [0, 2^{self.in_bits})")
        raw = b"".join(int(v).to_bytes(nb, "little") for v in vals)
        return torch.frombuffer(bytearray(raw), dtype=torch.uint8).reshape(len(vals), nb).to(device)

    @staticmethod
    def _dev(t: torch.Tensor) -> tuple[bool, int]:
        if t.device.type == "cuda":
            return True, t.device.index if t.device.index is not None else torch.cuda.current_device()
        return False, torch.cuda.current_device()

    @staticmethod
    def _need_cuda(what: str, *tensors: Optional[torch.Tensor]) -> None:
        """Device-pointer-only entry points: a host pointer handed to them would be dereferenced by a kernel."""
        for t in tensors:
            if t is not None and t.device.type != "cuda":
                raise RuntimeError(f"{what} takes CUDA tensors (got a {t.device.type} tensor); move the inputs to the GPU")

    @staticmethod
    def _need_rows(name: str, t: Optional[torch.Tensor], n: int) -> None:
        if t is not None and t.shape[0] != n:
            raise TypeError(f"{name} must have {n} rows, got {t.shape[0]}")

    @staticmethod
    def _check_out(out: Optional[torch.Tensor], shape, dtype, device) -> None:
        if out is None:
            return
        if tuple(out.shape) != tuple(shape) or out.dtype != dtype or out.device != device or not out.is_contiguous():
            raise TypeError(f"out must be a contiguous {dtype} tensor of shape {tuple(shape)} on {device}")

    @staticmethod
    def _stream(dev: int) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    # ---- Gen (dpf.cuh:93-159, dcf.cuh:108-194, half_tree_dpf.cuh:68-175, grotto_dcf.cuh:63-67) ---------------
    def gen(self, s0s: torch.Tensor, alphas: IntLike, betas: Optional[torch.Tensor] = None):
        s0s = s0s.contiguous()
        n = s0s.shape[0]
        on_gpu, dev = self._dev(s0s)
        al = self.in_tensor(alphas, s0s.device)
        self._need_rows("alphas", al, n)
        if self.scheme != "grotto":
            betas = betas.contiguous()
            self._need_rows("betas", betas, n)
        cws = torch.empty((n, self.ncw, 8), dtype=torch.int32, device=s0s.device)
        ocws = torch.empty((n, 4), dtype=torch.int32, device=s0s.device) if self.scheme == "halftree" else None
        h = self.handle(dev)
        if on_gpu:
            with torch.cuda.device(dev):
                L.check(L.lib.fssb200_gen(h, _ptr(s0s), _ptr(al), _ptr(betas), _ptr(cws), _ptr(ocws), n,
                                          self._stream(dev)), "fssb200_gen")
        else:
            self._ensure_host(dev)
            L.check(L.lib.fssb200_gen_host(h, _ptr(s0s), _ptr(al), _ptr(betas), _ptr(cws), _ptr(ocws), n),
                    "fssb200_gen_host")
        return (cws, ocws) if self.scheme == "halftree" else cws

    # ---- Eval (dpf.cuh:170-214, dcf.cuh:205-276, half_tree_dpf.cuh:187-231) ----------------------------------
    def eval(self, party: int, seeds: torch.Tensor, cws: torch.Tensor, xs: IntLike,
             ocws: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        seeds, cws = seeds.contiguous(), cws.contiguous()
        n = seeds.shape[0]
        on_gpu, dev = self._dev(seeds)
        x = self.in_tensor(xs, seeds.device)
        ocws = None if ocws is None else ocws.contiguous()
        self._need_rows("xs", x, n)
        self._need_rows("cws", cws, n)
        self._need_rows("ocws", ocws, n)
        self._check_out(out, (n, 4), torch.int32, seeds.device)
        ys = out if out is not None else torch.empty((n, 4), dtype=torch.int32, device=seeds.device)
        h = self.handle(dev)
        if on_gpu:
            with torch.cuda.device(dev):
                L.check(L.lib.fssb200_eval(h, party, _ptr(seeds), _ptr(cws), _ptr(ocws), _ptr(x), _ptr(ys), n,
                                           self._stream(dev)), "fssb200_eval")
        else:
            self._ensure_host(dev)
            L.check(L.lib.fssb200_eval_host(h, party, _ptr(seeds), _ptr(cws), _ptr(ocws), _ptr(x), _ptr(ys), n),
                    "fssb200_eval_host")
        return ys

    # ---- packed rows (compact key format, include/fssb200.h) ------------------------------------------------------
    def packed_row_bytes(self, dev: int = 0) -> int:
        return int(L.lib.fssb200_packed_row_bytes(self.handle(dev)))

    def host_pack_threads(self, dev: int = 0) -> int:
        """Threads the host entry point packs rows with (0: reference layout copied as it is)."""
        return int(L.lib.fssb200_ctx_host_pack_threads(self.handle(dev)))

    def pack_rows(self, cws: torch.Tensor, dev: int = 0) -> torch.Tensor:
        """Reference-layout keys (N, ncw, 8) int32 on the HOST -> packed rows (N, row_bytes) uint8 on the host."""
        if cws.device.type != "cpu":
            raise RuntimeError("pack_rows converts host arrays")
        cws = cws.contiguous()
        n = cws.shape[0]
        rb = self.packed_row_bytes(dev)
        if rb == 0:
            raise ValueError(f"scheme {self.scheme!r} has no packed key format")
        rows = torch.empty((n, rb), dtype=torch.uint8)
        L.check(L.lib.fssb200_pack_rows(self.handle(dev), _ptr(cws), _ptr(rows), n), "fssb200_pack_rows")
        return rows

    def eval_packed(self, party: int, seeds: torch.Tensor, rows: torch.Tensor, xs: IntLike,
                    ocws: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        seeds, rows = seeds.contiguous(), rows.contiguous()
        n = seeds.shape[0]
        on_gpu, dev = self._dev(seeds)
        if not on_gpu:
            raise RuntimeError("eval_packed takes device tensors (the host entry point packs internally)")
        x = self.in_tensor(xs, seeds.device)
        ocws = None if ocws is None else ocws.contiguous()
        self._need_cuda("eval_packed", rows, x, ocws)
        self._need_rows("xs", x, n)
        self._need_rows("rows", rows, n)
        self._check_out(out, (n, 4), torch.int32, seeds.device)
        ys = out if out is not None else torch.empty((n, 4), dtype=torch.int32, device=seeds.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_eval_packed(self.handle(dev), party, _ptr(seeds), _ptr(rows), _ptr(ocws), _ptr(x),
                                              _ptr(ys), n, self._stream(dev)), "fssb200_eval_packed")
        return ys

    # ---- EvalAll (dpf.cuh:232-303, half_tree_dpf.cuh:246-354, grotto_dcf.cuh:151-163) --------------------------
    def eval_all(self, party: int, seeds: torch.Tensor, cws: torch.Tensor, ocws: Optional[torch.Tensor] = None,
                 leaf_begin: int = 0, leaf_count: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        seeds, cws = seeds.contiguous(), cws.contiguous()
        n = seeds.shape[0]
        on_gpu, dev = self._dev(seeds)
        cnt = leaf_count or ((1 << self.in_bits) - leaf_begin)
        ocws = None if ocws is None else ocws.contiguous()
        self._need_rows("cws", cws, n)
        self._need_rows("ocws", ocws, n)
        if out is not None:
            want = (n, cnt) if self.scheme == "grotto" else (n, cnt, 4)
            self._check_out(out, want, torch.uint8 if self.scheme == "grotto" else torch.int32, seeds.device)
            ys = out
        elif self.scheme == "grotto":
            ys = torch.empty((n, cnt), dtype=torch.uint8, device=seeds.device)
        else:
            ys = torch.empty((n, cnt, 4), dtype=torch.int32, device=seeds.device)
        h = self.handle(dev)
        if on_gpu:
            with torch.cuda.device(dev):
                L.check(L.lib.fssb200_eval_all(h, party, _ptr(seeds), _ptr(cws), _ptr(ocws), _ptr(ys), n, leaf_begin,
                                               leaf_count, self._stream(dev)), "fssb200_eval_all")
        else:
            self._ensure_host(dev)
            L.check(L.lib.fssb200_eval_all_host(h, party, _ptr(seeds), _ptr(cws), _ptr(ocws), _ptr(ys), n,
                                                leaf_begin, leaf_count), "fssb200_eval_all_host")
        return ys

    # ---- level-major layout (point_eval_gpu.cuh:324-492) ----------------------------------------------------------
    def relayout(self, cws: torch.Tensor, out=None):
        """Key-major Cw[N][ncw] -> the level-major arrays (cw_s, cw_v, extra, out_cw); ``out`` reuses a previous result."""
        cws = cws.contiguous()
        self._need_cuda("relayout", cws)
        n, nb = cws.shape[0], self.in_bits
        _, dev = self._dev(cws)
        d = cws.device
        if out is not None:
            cw_s, cw_v, extra, out_cw = out
            self._check_out(cw_s, (nb, n, 4), torch.int32, d)
        else:
            cw_s = torch.empty((nb, n, 4), dtype=torch.int32, device=d)
            cw_v = torch.empty((nb, n, 4), dtype=torch.int32, device=d) if self.scheme == "dcf" else None
            extra = torch.empty(((nb + 31) // 32, n), dtype=torch.int32, device=d) if self.scheme != "dcf" else None
            out_cw = torch.empty((n, 4), dtype=torch.int32, device=d) if self.scheme not in ("halftree", "vdpf") else None
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_relayout(self.handle(dev), _ptr(cws), _ptr(cw_s), _ptr(cw_v), _ptr(extra),
                                           _ptr(out_cw), n, self._stream(dev)), "fssb200_relayout")
        return cw_s, cw_v, extra, out_cw

    def eval_levelmajor(self, party: int, seeds: torch.Tensor, layout, xs: IntLike,
                        ocws: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Point evaluation on the level-major (compact) key layout; CPU tensors go through
        ``fssb200_eval_levelmajor_host`` (strided chunk copies inside the library)."""
        cw_s, cw_v, extra, out_cw = layout
        seeds = seeds.contiguous()
        n = seeds.shape[0]
        on_gpu, dev = self._dev(seeds)
        x = self.in_tensor(xs, seeds.device)
        self._need_rows("xs", x, n)
        for t in (cw_s, cw_v, extra):
            if t is not None and (t.shape[1] != n or t.device != seeds.device):
                raise TypeError(f"level-major arrays must hold {n} keys on {seeds.device}")
        self._check_out(out, (n, 4), torch.int32, seeds.device)
        ys = out if out is not None else torch.empty((n, 4), dtype=torch.int32, device=seeds.device)
        if on_gpu:
            with torch.cuda.device(dev):
                L.check(L.lib.fssb200_eval_levelmajor(self.handle(dev), party, _ptr(seeds), _ptr(cw_s), _ptr(cw_v),
                                                      _ptr(extra), _ptr(out_cw), _ptr(ocws), _ptr(x), _ptr(ys), n,
                                                      self._stream(dev)), "fssb200_eval_levelmajor")
        else:
            self._ensure_host(dev)
            L.check(L.lib.fssb200_eval_levelmajor_host(self.handle(dev), party, _ptr(seeds), _ptr(cw_s), _ptr(cw_v),
                                                       _ptr(extra), _ptr(out_cw), _ptr(ocws), _ptr(x), _ptr(ys), n),
                    "fssb200_eval_levelmajor_host")
        return ys

    # ---- Grotto (grotto_dcf.cuh:94-135,174-238) ---------------------------------------------------------------------
    def grotto_expand(self, party: int, seeds: torch.Tensor, cws: torch.Tensor, leaf_begin: int = 0,
                      leaf_count: int = 0) -> torch.Tensor:
        seeds, cws = seeds.contiguous(), cws.contiguous()
        self._need_cuda("grotto_expand", seeds, cws)
        n = seeds.shape[0]
        self._need_rows("cws", cws, n)
        _, dev = self._dev(seeds)
        cnt = leaf_count or ((1 << self.in_bits) - leaf_begin)
        t = torch.empty((n, cnt), dtype=torch.uint8, device=seeds.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_grotto_expand(self.handle(dev), party, _ptr(seeds), _ptr(cws), _ptr(t), n,
                                                leaf_begin, leaf_count, self._stream(dev)), "fssb200_grotto_expand")
        return t

    def grotto_preprocess(self, party: int, seeds: torch.Tensor, cws: torch.Tensor) -> torch.Tensor:
        seeds, cws = seeds.contiguous(), cws.contiguous()
        self._need_cuda("grotto_preprocess", seeds, cws)
        n = seeds.shape[0]
        self._need_rows("cws", cws, n)
        _, dev = self._dev(seeds)
        pt = torch.empty((n, (2 << self.in_bits) - 1), dtype=torch.uint8, device=seeds.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_grotto_preprocess(self.handle(dev), party, _ptr(seeds), _ptr(cws), _ptr(pt), n,
                                                    self._stream(dev)), "fssb200_grotto_preprocess")
        return pt

    def grotto_lookup(self, pt: torch.Tensor, xs: IntLike) -> torch.Tensor:
        pt = pt.contiguous()
        self._need_cuda("grotto_lookup", pt)
        n = pt.shape[0]
        _, dev = self._dev(pt)
        x = self.in_tensor(xs, pt.device)
        self._need_rows("xs", x, n)
        if pt.dim() != 2 or pt.shape[1] != (2 << self.in_bits) - 1 or pt.dtype != torch.uint8:
            raise TypeError(f"pt must be a (N, {(2 << self.in_bits) - 1}) uint8 tensor")
        ys = torch.empty((n,), dtype=torch.uint8, device=pt.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_grotto_eval(self.handle(dev), _ptr(pt), _ptr(x), _ptr(ys), n, self._stream(dev)),
                    "fssb200_grotto_eval")
        return ys

    def grotto_walk(self, party: int, seeds: torch.Tensor, cws: torch.Tensor, xs: IntLike,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """O(n) Grotto point evaluation (fssb200_grotto_eval_walk): (N,) uint8 shares with
        share0 ^ share1 = 1[alpha <= x]; reconstruction-equal to GrottoDcf::Eval (grotto_dcf.cuh:116-135), per-party
        bits differ.  Works for any in_bits (no 2N-1-byte parity tree)."""
        seeds, cws = seeds.contiguous(), cws.contiguous()
        self._need_cuda("grotto_walk", seeds, cws)
        n = seeds.shape[0]
        _, dev = self._dev(seeds)
        x = self.in_tensor(xs, seeds.device)
        self._need_rows("xs", x, n)
        self._need_rows("cws", cws, n)
        self._check_out(out, (n,), torch.uint8, seeds.device)
        ys = out if out is not None else torch.empty((n,), dtype=torch.uint8, device=seeds.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_grotto_eval_walk(self.handle(dev), party, _ptr(seeds), _ptr(cws), _ptr(x), _ptr(ys), n,
                                                   self._stream(dev)), "fssb200_grotto_eval_walk")
        return ys

    # ---- VDPF (vdpf.cuh) ---------------------------------------------------------------------------------------------
    def vdpf_gen(self, s0s: torch.Tensor, alphas: IntLike, betas: torch.Tensor):
        """Vdpf::Gen (vdpf.cuh:97-177) -> cws (N,n,8), cs (N,4,4), ocws (N,4), status (N,) int32
        (1 = t0 == t1: resample that key's seeds; its ocw row is zero)."""
        s0s, betas = s0s.contiguous(), betas.contiguous()
        n = s0s.shape[0]
        on_gpu, dev = self._dev(s0s)
        d = s0s.device
        al = self.in_tensor(alphas, d)
        self._need_rows("alphas", al, n)
        self._need_rows("betas", betas, n)
        cws = torch.empty((n, self.ncw, 8), dtype=torch.int32, device=d)
        cs = torch.empty((n, 4, 4), dtype=torch.int32, device=d)
        ocws = torch.zeros((n, 4), dtype=torch.int32, device=d)
        status = torch.empty((n,), dtype=torch.int32, device=d)
        h = self.handle(dev)
        if on_gpu:
            with torch.cuda.device(dev):
                L.check(L.lib.fssb200_vdpf_gen(h, _ptr(s0s), _ptr(al), _ptr(betas), _ptr(cws), _ptr(cs), _ptr(ocws),
                                               _ptr(status), n, self._stream(dev)), "fssb200_vdpf_gen")
        else:
            self._ensure_host(dev)
            L.check(L.lib.fssb200_vdpf_gen_host(h, _ptr(s0s), _ptr(al), _ptr(betas), _ptr(cws), _ptr(cs), _ptr(ocws),
                                                _ptr(status), n), "fssb200_vdpf_gen_host")
        return cws, cs, ocws, status

    def vdpf_eval(self, party: int, seeds: torch.Tensor, cws: torch.Tensor, cs: torch.Tensor, ocws: torch.Tensor,
                  xs: IntLike, layout=None):
        """Vdpf::Eval (vdpf.cuh:191-243) -> ys (N,4), pi_tildes (N,4,4).  ``layout`` = (cw_s, extra) from
        ``relayout`` selects the level-major kernel (point_eval_gpu.cuh:514-527)."""
        seeds, cs, ocws = seeds.contiguous(), cs.contiguous(), ocws.contiguous()
        n = seeds.shape[0]
        on_gpu, dev = self._dev(seeds)
        d = seeds.device
        x = self.in_tensor(xs, d)
        self._need_rows("xs", x, n)
        self._need_rows("cs", cs, n)
        self._need_rows("ocws", ocws, n)
        ys = torch.empty((n, 4), dtype=torch.int32, device=d)
        pis = torch.empty((n, 4, 4), dtype=torch.int32, device=d)
        h = self.handle(dev)
        if layout is not None:
            cw_s, extra = layout
            self._need_cuda("vdpf_eval(layout=...)", seeds, cw_s, extra, cs, ocws)
            with torch.cuda.device(dev):
                L.check(L.lib.fssb200_vdpf_eval_levelmajor(h, party, _ptr(seeds), _ptr(cw_s), _ptr(extra), _ptr(cs),
                                                           _ptr(ocws), _ptr(x), _ptr(ys), _ptr(pis), n,
                                                           self._stream(dev)), "fssb200_vdpf_eval_levelmajor")
            return ys, pis
        cws = cws.contiguous()
        if on_gpu:
            with torch.cuda.device(dev):
                L.check(L.lib.fssb200_vdpf_eval(h, party, _ptr(seeds), _ptr(cws), _ptr(cs), _ptr(ocws), _ptr(x),
                                                _ptr(ys), _ptr(pis), n, self._stream(dev)), "fssb200_vdpf_eval")
        else:
            self._ensure_host(dev)
            L.check(L.lib.fssb200_vdpf_eval_host(h, party, _ptr(seeds), _ptr(cws), _ptr(cs), _ptr(ocws), _ptr(x),
                                                 _ptr(ys), _ptr(pis), n), "fssb200_vdpf_eval_host")
        return ys, pis

    def vdpf_prove(self, pi_tildes: torch.Tensor, cs: torch.Tensor) -> torch.Tensor:
        """Vdpf::Prove (vdpf.cuh:254-264): pi_tildes (N,m,4,4), cs (N,4,4) -> proofs (N,4,4)."""
        pi_tildes, cs = pi_tildes.contiguous(), cs.contiguous()
        self._need_cuda("vdpf_prove", pi_tildes, cs)
        n, m = cs.shape[0], pi_tildes.shape[1]
        self._need_rows("pi_tildes", pi_tildes, n)
        _, dev = self._dev(cs)
        pis = torch.empty((n, 4, 4), dtype=torch.int32, device=cs.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_vdpf_prove(self.handle(dev), _ptr(pi_tildes), _ptr(cs), m, _ptr(pis), n,
                                             self._stream(dev)), "fssb200_vdpf_prove")
        return pis

    def vdpf_eval_all(self, party: int, seeds: torch.Tensor, cws: torch.Tensor, cs: torch.Tensor,
                      ocws: torch.Tensor):
        """Vdpf::EvalAll (vdpf.cuh:294-342) -> ys (N,2^n,4), proofs (N,4,4)."""
        seeds, cws, cs, ocws = seeds.contiguous(), cws.contiguous(), cs.contiguous(), ocws.contiguous()
        self._need_cuda("vdpf_eval_all", seeds, cws, cs, ocws)
        n = seeds.shape[0]
        _, dev = self._dev(seeds)
        ys = torch.empty((n, 1 << self.in_bits, 4), dtype=torch.int32, device=seeds.device)
        pis = torch.empty((n, 4, 4), dtype=torch.int32, device=seeds.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_vdpf_eval_all(self.handle(dev), party, _ptr(seeds), _ptr(cws), _ptr(cs), _ptr(ocws),
                                                _ptr(ys), _ptr(pis), n, self._stream(dev)), "fssb200_vdpf_eval_all")
        return ys, pis

    @staticmethod
    def vdpf_verify(pi0: torch.Tensor, pi1: torch.Tensor) -> torch.Tensor:
        """Vdpf::Verify (vdpf.cuh:271-276): per-key equality of two proofs (N,4,4) -> bool (N,)."""
        return (pi0 == pi1).flatten(1).all(dim=1)

    def hash(self, which: int, msgs: torch.Tensor) -> torch.Tensor:
        """Blake3 plugin known-answer hook: which 0 = XorHash (N,2,4)->(N,4,4), 1 = Hash (N,4,4)->(N,2,4)."""
        msgs = msgs.contiguous()
        self._need_cuda("hash", msgs)
        n = msgs.shape[0]
        _, dev = self._dev(msgs)
        out = torch.empty((n, 4 if which == 0 else 2, 4), dtype=torch.int32, device=msgs.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_hash(self.handle(dev), which, _ptr(msgs), _ptr(out), n, self._stream(dev)),
                    "fssb200_hash")
        return out

    # ---- PRG known-answer hook -----------------------------------------------------------------------------------------
    def prg_gen(self, seeds: torch.Tensor, mul: int) -> torch.Tensor:
        seeds = seeds.contiguous()
        self._need_cuda("prg_gen", seeds)
        n = seeds.shape[0]
        _, dev = self._dev(seeds)
        out = torch.empty((n, mul, 4), dtype=torch.int32, device=seeds.device)
        with torch.cuda.device(dev):
            L.check(L.lib.fssb200_prg_gen(self.handle(dev), _ptr(seeds), _ptr(out), mul, n, self._stream(dev)),
                    "fssb200_prg_gen")
        return out


def microbench(kind: int, device: int = 0) -> float:
    """Issue-rate microbenchmark (ops/s): 0 LOP3, 1 IMAD, 2 LOP3+IMAD, 3 conflict-free LDS.32, 4 PRMT."""
    v = C.c_double(0.0)
    L.check(L.lib.fssb200_microbench(device, kind, C.byref(v)), "fssb200_microbench")
    return float(v.value)
