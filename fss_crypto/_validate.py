"""Alias of fss_b200._validate under the reference's module path (fss_crypto/_validate.py)."""
from fss_b200._validate import *  # noqa: F401,F403
from fss_b200._validate import (validate_alpha, validate_beta, validate_cpu_only, validate_cws,  # noqa: F401
                                validate_device_match, validate_domain_value, validate_group, validate_in_bits,
                                validate_party, validate_pred, validate_prg, validate_s0, validate_s0s)
