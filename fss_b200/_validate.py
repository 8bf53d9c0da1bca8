"""Argument validation of the fss_crypto-compatible front-end.

The reference's validators (fss_crypto/_validate.py:16-108) are a public contract: its
test/test_validation.py imports them by name and pins exception types and message formats.
Here they are generated from two rule tables (allowed spellings; int32 tensor shapes) so that
the single-key shapes of the reference and the batched ``(N, ...)`` shapes this package adds
go through the same two checkers.
"""
from __future__ import annotations

from numbers import Integral

import torch

# argument -> allowed spellings (reference _validate.py:7-13)
_CHOICES = {
    "group": ("bytes", "uint"),
    "prg": ("chacha", "aes128_mmo"),
    "pred": ("lt", "gt"),
    "party": (0, 1),
}
# scheme -> PRGs it can be instantiated with (reference _validate.py:9-12)
_PRGS_OF = {"dpf": _CHOICES["prg"], "dcf": _CHOICES["prg"]}
# argument -> shape of its int32 tensor; in_bits-dependent entries are callables
_INT32_SHAPES = {
    "s0": (4,),
    "s0s": (2, 4),
    "beta": (4,),
    "cws": lambda in_bits: (in_bits + 1, 8),
}


def _require_choice(name: str, value, allowed, shown=None) -> None:
    if value not in allowed:
        raise ValueError(f"{name} must be {shown or f'one of {allowed}'}, got {value if name == 'party' else repr(value)}")


def _require_int32(name: str, t: torch.Tensor, shape, shown=None) -> None:
    """``t`` is int32 and ``t.shape`` equals ``shape`` where ``shape`` has an int, anything where it has None."""
    ok = t.dtype == torch.int32 and t.dim() == len(shape) and all(
        want is None or want == got for want, got in zip(shape, t.shape))
    if not ok:
        shown = shown if shown is not None else tuple(shape)
        raise TypeError(f"{name} must be a {shown} int32 tensor, got shape {tuple(t.shape)} dtype {t.dtype}")


def validate_in_bits(in_bits: int) -> None:
    if in_bits < 1 or in_bits > 128:
        raise ValueError(f"in_bits must be between 1 and 128, got {in_bits}")


def validate_group(group: str) -> None:
    _require_choice("group", group, _CHOICES["group"])


def validate_prg(prg: str, scheme: str) -> None:
    _require_choice("scheme", scheme, _PRGS_OF, f"one of {tuple(_PRGS_OF)}")
    _require_choice("prg", prg, _PRGS_OF[scheme])


def validate_pred(pred: str) -> None:
    _require_choice("pred", pred, _CHOICES["pred"])


def validate_party(party: int) -> None:
    _require_choice("party", party, _CHOICES["party"], "0 or 1")


def validate_s0(s0: torch.Tensor) -> None:
    _require_int32("s0", s0, _INT32_SHAPES["s0"])


def validate_s0s(s0s: torch.Tensor) -> None:
    _require_int32("s0s", s0s, _INT32_SHAPES["s0s"])


def validate_beta(beta: torch.Tensor) -> None:
    _require_int32("beta", beta, _INT32_SHAPES["beta"])


def validate_cws(cws: torch.Tensor, in_bits: int) -> None:
    _require_int32("cws", cws, _INT32_SHAPES["cws"](in_bits))


def validate_batched(name: str, t: torch.Tensor, tail: tuple) -> int:
    """Batched extension (not in the reference): ``t`` must be int32 with shape ``(N, *tail)``; returns N."""
    tail = tuple(tail)
    _require_int32(name, t, (None,) + tail, ("N",) + tail)
    return int(t.shape[0])


def validate_domain_value(name: str, value: int, in_bits: int) -> None:
    """A point of the input domain: a real integer (bool is rejected) below 2^in_bits."""
    if isinstance(value, bool) or not isinstance(value, Integral):
        raise TypeError(f"{name} must be an integer, got {type(value).__name__}")
    if not 0 <= value < (1 << in_bits):
        raise ValueError(f"{name} must be in [0, 2^{in_bits}), got {value}")


def validate_alpha(alpha: int, in_bits: int) -> None:
    validate_domain_value("alpha", alpha, in_bits)


def validate_device_match(*tensors: torch.Tensor) -> None:
    seen = sorted({str(t.device) for t in tensors})
    if len(seen) > 1:
        raise RuntimeError(
            f"expected all tensors to be on the same device, but found at least two devices, {', '.join(seen)}!")


def validate_cpu_only(*tensors: torch.Tensor, fn_name: str = "") -> None:
    head = f"{fn_name} expects" if fn_name else "expected"
    for t in tensors:
        if t.device.type != "cpu":
            raise RuntimeError(f"{head} all tensors to be on cpu, but found tensor on {t.device}")
