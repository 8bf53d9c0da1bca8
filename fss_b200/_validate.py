"""Argument validation of the fss_crypto-compatible front-end.

Mirrors the reference's validators (fss_crypto/_validate.py:16-108) -- same function names,
exception types and message formats, because test/test_validation.py pins them -- extended
with the batched shapes ``(N, ...)`` this package adds.
"""
from __future__ import annotations

from numbers import Integral

import torch

_VALID_GROUPS = ("bytes", "uint")
_VALID_PRGS = ("chacha", "aes128_mmo")
_VALID_PRGS_BY_SCHEME = {"dpf": _VALID_PRGS, "dcf": _VALID_PRGS}
_VALID_PREDS = ("lt", "gt")


def validate_in_bits(in_bits: int) -> None:
    if not (1 <= in_bits <= 128):
        raise ValueError(f"in_bits must be between 1 and 128, got {in_bits}")


def validate_group(group: str) -> None:
    if group not in _VALID_GROUPS:
        raise ValueError(f"group must be one of {_VALID_GROUPS}, got {group!r}")


def validate_prg(prg: str, scheme: str) -> None:
    valid = _VALID_PRGS_BY_SCHEME.get(scheme)
    if valid is None:
        raise ValueError(f"scheme must be one of {tuple(_VALID_PRGS_BY_SCHEME)}, got {scheme!r}")
    if prg not in valid:
        raise ValueError(f"prg must be one of {valid}, got {prg!r}")


def validate_pred(pred: str) -> None:
    if pred not in _VALID_PREDS:
        raise ValueError(f"pred must be one of {_VALID_PREDS}, got {pred!r}")


def validate_party(party: int) -> None:
    if party not in (0, 1):
        raise ValueError(f"party must be 0 or 1, got {party}")


def _shape_err(name: str, want, t: torch.Tensor) -> TypeError:
    return TypeError(f"{name} must be a {want} int32 tensor, got shape {tuple(t.shape)} dtype {t.dtype}")


def validate_s0(s0: torch.Tensor) -> None:
    if s0.shape != (4,) or s0.dtype != torch.int32:
        raise _shape_err("s0", "(4,)", s0)


def validate_s0s(s0s: torch.Tensor) -> None:
    if s0s.shape != (2, 4) or s0s.dtype != torch.int32:
        raise _shape_err("s0s", "(2, 4)", s0s)


def validate_beta(beta: torch.Tensor) -> None:
    if beta.shape != (4,) or beta.dtype != torch.int32:
        raise _shape_err("beta", "(4,)", beta)


def validate_cws(cws: torch.Tensor, in_bits: int) -> None:
    expected = (in_bits + 1, 8)
    if cws.shape != expected or cws.dtype != torch.int32:
        raise _shape_err("cws", expected, cws)


def validate_domain_value(name: str, value: int, in_bits: int) -> None:
    if isinstance(value, bool) or not isinstance(value, Integral):
        raise TypeError(f"{name} must be an integer, got {type(value).__name__}")
    if value < 0 or value >= (1 << in_bits):
        raise ValueError(f"{name} must be in [0, 2^{in_bits}), got {value}")


def validate_alpha(alpha: int, in_bits: int) -> None:
    validate_domain_value("alpha", alpha, in_bits)


def validate_device_match(*tensors: torch.Tensor) -> None:
    devices = {t.device for t in tensors}
    if len(devices) > 1:
        dev_list = ", ".join(str(d) for d in sorted(devices, key=str))
        raise RuntimeError(
            f"expected all tensors to be on the same device, but found at least two devices, {dev_list}!")


def validate_cpu_only(*tensors: torch.Tensor, fn_name: str = "") -> None:
    for t in tensors:
        if t.device.type != "cpu":
            prefix = f"{fn_name} expects" if fn_name else "expected"
            raise RuntimeError(f"{prefix} all tensors to be on cpu, but found tensor on {t.device}")


# ---- batched extensions (not in the reference) ---------------------------------------------------

def validate_batched(name: str, t: torch.Tensor, tail: tuple) -> int:
    """``t`` must be int32 with shape ``(N, *tail)``; returns N."""
    if t.dim() != len(tail) + 1 or tuple(t.shape[1:]) != tuple(tail) or t.dtype != torch.int32:
        raise _shape_err(name, ("N",) + tuple(tail), t)
    return int(t.shape[0])
