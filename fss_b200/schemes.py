"""fss_crypto-compatible scheme classes over the B200 evaluator.

``Dpf`` and ``Dcf`` keep the reference binding's constructors and single-key methods
(fss_crypto/dpf.py:23-109, fss_crypto/dcf.py): same argument names, shapes, dtypes, exception
types and messages, so the reference's own test/ directory runs against them unchanged.  Added on
top: a leading batch dimension on every tensor argument, CUDA tensors everywhere (the reference
evaluates one key with a ``<<<1,1>>>`` kernel, fss_crypto/_csrc/dpf_binding_impl.cuh:76-82),
explicit PRG key material (the reference's AES path does not compile and its ChaCha nonce is
process-random), and the Half-Tree / Grotto schemes the reference binding
does not expose.

Every method evaluates on the GPU through libfssb200.so.  gen / eval / eval_all accept CPU tensors as
in the reference (results come back on the inputs' device) and travel through the library's
host-buffer entry points; the Grotto preprocess / lookup / walk, relayout and VDPF prove / eval_all
methods take CUDA tensors only and raise RuntimeError otherwise.  There is no CPU evaluation path.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._validate import (validate_alpha, validate_batched, validate_beta, validate_cws, validate_device_match,
                        validate_domain_value, validate_group, validate_in_bits, validate_party, validate_pred,
                        validate_prg, validate_s0, validate_s0s)
from .context import Context, IntLike


def _uint_group(in_bits: int) -> str:
    """``group="uint"`` -> Uint<uint32_t> / Uint<uint64_t> / Uint<__uint128_t, 2^127> by in_bits
    (fss_crypto/_jit.py:76-87)."""
    return "u32" if in_bits <= 32 else ("u64" if in_bits <= 64 else "u128")


def _prg_key(prg: str, aes_keys, nonce) -> Optional[bytes]:
    if prg == "aes128_mmo" and aes_keys is not None:
        t = torch.as_tensor(aes_keys, dtype=torch.uint8).reshape(-1)
        if t.numel() % 16 or t.numel() > 64:
            raise TypeError("aes_keys must be a (mul, 16) uint8 tensor")
        return bytes(t.tolist())
    if prg == "chacha" and nonce is not None:
        t = torch.as_tensor(nonce, dtype=torch.int32).reshape(-1)
        if t.numel() != 2:
            raise TypeError("nonce must be a (2,) int32 tensor")
        return t.numpy().tobytes()
    return None


class _PointScheme:
    """Shared gen / eval / eval_all plumbing of Dpf and Dcf."""
    _scheme = "dpf"

    def _init(self, in_bits, group, prg, pred, aes_keys, nonce, hash_key=None):
        self.in_bits, self.group, self.prg = in_bits, group, prg
        g = "bytes" if group == "bytes" else _uint_group(in_bits)
        self._ctx = Context(self._scheme, in_bits, g, 0, prg, pred, _prg_key(prg, aes_keys, nonce), hash_key)

    @property
    def context(self) -> Context:
        return self._ctx

    # -- gen --------------------------------------------------------------------------------------------
    def gen(self, s0s: torch.Tensor, alpha: IntLike, beta: torch.Tensor) -> torch.Tensor:
        """Generate keys.  Single: s0s (2,4), alpha int, beta (4,) -> cws (in_bits+1, 8).
        Batched: s0s (N,2,4), alpha (N,) ints / tensor, beta (N,4) -> cws (N, in_bits+1, 8)."""
        if s0s.dim() == 2:
            validate_s0s(s0s)
            validate_alpha(alpha, self.in_bits)
            validate_beta(beta)
            validate_device_match(s0s, beta)
            return self._ctx.gen(s0s.unsqueeze(0), [alpha], beta.unsqueeze(0))[0]
        n = validate_batched("s0s", s0s, (2, 4))
        if validate_batched("beta", beta, (4,)) != n:
            raise TypeError(f"beta must have {n} rows")
        validate_device_match(s0s, beta)
        return self._ctx.gen(s0s, alpha, beta)

    # -- eval -------------------------------------------------------------------------------------------
    def eval(self, party: int, s0: torch.Tensor, cws: torch.Tensor, x: IntLike) -> torch.Tensor:
        """Evaluate.  Single: s0 (4,), cws (in_bits+1, 8), x int -> (4,).
        Batched: s0 (N,4), cws (N, in_bits+1, 8), x (N,) -> (N,4).  Output on the inputs' device."""
        validate_party(party)
        if s0.dim() == 1:
            validate_s0(s0)
            validate_cws(cws, self.in_bits)
            validate_device_match(s0, cws)
            validate_domain_value("x", x, self.in_bits)
            return self._ctx.eval(party, s0.unsqueeze(0), cws.unsqueeze(0), [x])[0]
        n = validate_batched("s0", s0, (4,))
        if validate_batched("cws", cws, (self.in_bits + 1, 8)) != n:
            raise TypeError(f"cws must have {n} rows")
        validate_device_match(s0, cws)
        return self._ctx.eval(party, s0, cws, x)

    # -- eval_all ---------------------------------------------------------------------------------------
    def eval_all(self, party: int, s0: torch.Tensor, cws: torch.Tensor) -> torch.Tensor:
        """Full-domain evaluation.  Single: -> (2^in_bits, 4); batched: -> (N, 2^in_bits, 4)."""
        validate_party(party)
        if s0.dim() == 1:
            validate_s0(s0)
            validate_cws(cws, self.in_bits)
            validate_device_match(s0, cws)
            return self._ctx.eval_all(party, s0.unsqueeze(0), cws.unsqueeze(0))[0]
        n = validate_batched("s0", s0, (4,))
        if validate_batched("cws", cws, (self.in_bits + 1, 8)) != n:
            raise TypeError(f"cws must have {n} rows")
        validate_device_match(s0, cws)
        return self._ctx.eval_all(party, s0, cws)


class Dpf(_PointScheme):
    """2-party Distributed Point Function (fss_crypto/dpf.py:23-41).

    Args:
        in_bits: Input domain bit size (1..128).
        group: Output group type, "bytes" or "uint".
        prg: PRG type, "chacha" or "aes128_mmo".
        aes_keys: optional (2, 16) uint8 AES user keys (default: the reference samples' keys).
        nonce: optional (2,) int32 ChaCha nonce (default {0x12345678, 0x9abcdef0}).
    """
    _scheme = "dpf"

    def __init__(self, in_bits: int, group: str = "bytes", prg: str = "chacha", *, aes_keys=None, nonce=None):
        validate_in_bits(in_bits)
        validate_group(group)
        validate_prg(prg, "dpf")
        self._init(in_bits, group, prg, "lt", aes_keys, nonce)


class Dcf(_PointScheme):
    """2-party Distributed Comparison Function (fss_crypto/dcf.py:23-47).

    Args:
        in_bits: Input domain bit size (1..128).
        group: Output group type, "bytes" or "uint".
        prg: PRG type, "chacha" or "aes128_mmo".
        pred: Comparison predicate, "lt" (less-than) or "gt" (greater-than).
    """
    _scheme = "dcf"

    def __init__(self, in_bits: int, group: str = "bytes", prg: str = "chacha", pred: str = "lt", *, aes_keys=None,
                 nonce=None):
        validate_in_bits(in_bits)
        validate_group(group)
        validate_prg(prg, "dcf")
        validate_pred(pred)
        self.pred = pred
        self._init(in_bits, group, prg, pred, aes_keys, nonce)


class HalfTreeDpf:
    """Half-Tree DPF (half_tree_dpf.cuh:39-355); batched tensors only.

    gen(s0s (N,2,4), alpha, beta (N,4)) -> (cws (N, in_bits, 8), ocws (N,4));
    eval(party, s0 (N,4), cws, ocws, x) -> (N,4); eval_all(...) -> (N, 2^in_bits, 4)."""

    def __init__(self, in_bits: int, group: str = "bytes", prg: str = "chacha", *, hash_key=None, aes_keys=None,
                 nonce=None):
        validate_in_bits(in_bits)
        validate_group(group)
        validate_prg(prg, "dpf")
        self.in_bits, self.group, self.prg = in_bits, group, prg
        hk = None if hash_key is None else torch.as_tensor(hash_key, dtype=torch.int32).numpy().tobytes()
        g = "bytes" if group == "bytes" else _uint_group(in_bits)
        self._ctx = Context("halftree", in_bits, g, 0, prg, "lt", _prg_key(prg, aes_keys, nonce), hk)

    @property
    def context(self) -> Context:
        return self._ctx

    def gen(self, s0s, alpha, beta):
        validate_batched("s0s", s0s, (2, 4))
        validate_batched("beta", beta, (4,))
        return self._ctx.gen(s0s, alpha, beta)

    def eval(self, party, s0, cws, ocws, x):
        validate_party(party)
        validate_batched("s0", s0, (4,))
        validate_batched("cws", cws, (self.in_bits, 8))
        validate_batched("ocws", ocws, (4,))
        return self._ctx.eval(party, s0, cws, x, ocws)

    def eval_all(self, party, s0, cws, ocws):
        validate_party(party)
        validate_batched("s0", s0, (4,))
        validate_batched("cws", cws, (self.in_bits, 8))
        return self._ctx.eval_all(party, s0, cws, ocws)


class GrottoDcf:
    """Grotto DCF over F2 (grotto_dcf.cuh:45-239); batched tensors only.

    gen(s0s (N,2,4), alpha) -> cws (N, in_bits+1, 8); eval_all(party, s0, cws) -> (N, 2^in_bits) uint8 shares of
    1[alpha <= x]; preprocess(...) -> parity trees (N, 2^(in_bits+1) - 1) uint8; eval(pt, x) -> (N,) uint8."""

    def __init__(self, in_bits: int, prg: str = "chacha", *, aes_keys=None, nonce=None):
        validate_in_bits(in_bits)
        validate_prg(prg, "dpf")
        self.in_bits, self.prg = in_bits, prg
        self._ctx = Context("grotto", in_bits, "bytes", 0, prg, "lt", _prg_key(prg, aes_keys, nonce), None)

    @property
    def context(self) -> Context:
        return self._ctx

    def gen(self, s0s, alpha):
        validate_batched("s0s", s0s, (2, 4))
        return self._ctx.gen(s0s, alpha, None)

    def eval_all(self, party, s0, cws):
        validate_party(party)
        return self._ctx.eval_all(party, s0, cws)

    def preprocess(self, party, s0, cws):
        validate_party(party)
        return self._ctx.grotto_preprocess(party, s0, cws)

    def eval(self, pt, x):
        return self._ctx.grotto_lookup(pt, x)

    def eval_walk(self, party, s0, cws, x):
        """O(n) point evaluation without a parity tree, any in_bits (e.g. BASELINE configs[4]: n = 32, 2^20 keys):
        (N,) uint8 shares, share0 ^ share1 = 1[alpha <= x] as for eval(); the per-party bit is not the reference's
        (include/fssb200.h: fssb200_grotto_eval_walk)."""
        validate_party(party)
        validate_batched("s0", s0, (4,))
        validate_batched("cws", cws, (self.in_bits + 1, 8))
        return self._ctx.grotto_walk(party, s0, cws, x)
