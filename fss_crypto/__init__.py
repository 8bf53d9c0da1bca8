"""``import fss_crypto`` (the reference binding's name) resolves to the B200 evaluator."""
from fss_b200.schemes import Dcf, Dpf

__all__ = ["Dcf", "Dpf"]
