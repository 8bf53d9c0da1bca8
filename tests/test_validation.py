"""Validator contract of the fss_crypto-compatible front-end: exception types and message formats the
reference pins in test/test_validation.py (accept / reject cases restated, not copied)."""
import pytest
import torch

from fss_crypto._validate import (validate_alpha, validate_beta, validate_cpu_only, validate_cws,
                                  validate_device_match, validate_domain_value, validate_group, validate_in_bits,
                                  validate_party, validate_pred, validate_prg, validate_s0, validate_s0s)
from fss_b200._validate import validate_batched


def i32(*shape):
    return torch.zeros(shape, dtype=torch.int32)


def test_in_bits():
    for ok in (1, 64, 128):
        validate_in_bits(ok)
    for bad in (0, 129, -5):
        with pytest.raises(ValueError, match="in_bits must be between 1 and 128"):
            validate_in_bits(bad)


def test_group_prg_pred_party():
    validate_group("bytes"); validate_group("uint")
    with pytest.raises(ValueError, match="group must be one of"):
        validate_group("invalid")
    for s in ("dpf", "dcf"):
        validate_prg("chacha", s); validate_prg("aes128_mmo", s)
    with pytest.raises(ValueError, match="prg must be one of"):
        validate_prg("invalid", "dpf")
    with pytest.raises(ValueError, match="scheme must be one of"):
        validate_prg("chacha", "invalid")
    validate_pred("lt"); validate_pred("gt")
    with pytest.raises(ValueError, match="pred must be one of"):
        validate_pred("le")
    validate_party(0); validate_party(1)
    for bad in (2, -1):
        with pytest.raises(ValueError, match="party must be 0 or 1"):
            validate_party(bad)


def test_tensor_shapes():
    validate_s0(i32(4)); validate_s0s(i32(2, 4)); validate_beta(i32(4)); validate_cws(i32(17, 8), 16)
    with pytest.raises(TypeError, match=r"s0 must be a \(4,\) int32 tensor"):
        validate_s0(i32(5))
    with pytest.raises(TypeError, match=r"s0 must be a \(4,\) int32 tensor"):
        validate_s0(torch.zeros(4, dtype=torch.int64))
    with pytest.raises(TypeError, match=r"s0s must be a \(2, 4\) int32 tensor"):
        validate_s0s(i32(4))
    with pytest.raises(TypeError, match=r"beta must be a \(4,\) int32 tensor"):
        validate_beta(i32(4, 1))
    with pytest.raises(TypeError, match=r"cws must be a \(17, 8\) int32 tensor"):
        validate_cws(i32(16, 8), 16)
    with pytest.raises(TypeError, match=r"cws must be a \(17, 8\) int32 tensor"):
        validate_cws(torch.zeros(17, 8), 16)
    assert validate_batched("s0", i32(9, 4), (4,)) == 9
    with pytest.raises(TypeError, match="s0 must be a"):
        validate_batched("s0", i32(9, 5), (4,))


def test_domain_values():
    validate_domain_value("x", 0, 16); validate_domain_value("x", 2 ** 16 - 1, 16); validate_alpha(2 ** 127, 128)
    with pytest.raises(ValueError, match=r"x must be in \[0, 2\^16\)"):
        validate_domain_value("x", 2 ** 16, 16)
    with pytest.raises(ValueError, match=r"alpha must be in \[0, 2\^8\)"):
        validate_alpha(-1, 8)
    with pytest.raises(TypeError, match="x must be an integer, got float"):
        validate_domain_value("x", 1.5, 16)
    with pytest.raises(TypeError, match="x must be an integer, got bool"):
        validate_domain_value("x", True, 16)


def test_device_rules():
    validate_device_match(i32(4), i32(17, 8))
    validate_cpu_only(i32(4), fn_name="gen")
    meta = torch.zeros(4, dtype=torch.int32, device="meta")
    with pytest.raises(RuntimeError, match="expected all tensors to be on the same device"):
        validate_device_match(i32(4), meta)
    with pytest.raises(RuntimeError, match="gen expects all tensors to be on cpu"):
        validate_cpu_only(meta, fn_name="gen")
    with pytest.raises(RuntimeError, match="expected all tensors to be on cpu"):
        validate_cpu_only(meta)


def test_constructors_validate_before_touching_the_gpu():
    import fss_crypto
    with pytest.raises(ValueError, match="in_bits must be between"):
        fss_crypto.Dpf(0)
    with pytest.raises(ValueError, match="group must be one of"):
        fss_crypto.Dpf(16, group="int")
    with pytest.raises(ValueError, match="prg must be one of"):
        fss_crypto.Dcf(16, prg="sha")
    with pytest.raises(ValueError, match="pred must be one of"):
        fss_crypto.Dcf(16, pred="ge")
    d = fss_crypto.Dpf(in_bits=16, group="bytes", prg="chacha")   # lazy: no device needed to construct
    assert (d.in_bits, d.group, d.prg) == (16, "bytes", "chacha")
    with pytest.raises(ValueError, match="x must be"):
        d.eval(party=0, s0=i32(4), cws=i32(17, 8), x=2 ** 16)
    with pytest.raises(TypeError, match="cws must be a"):
        d.eval(party=0, s0=i32(4), cws=i32(16, 8), x=1)
