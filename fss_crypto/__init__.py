"""Drop-in alias: ``import fss_crypto`` resolves to the B200 evaluator (fss_b200), so code written
against the reference binding (fss_crypto/__init__.py:3-6) runs unchanged."""
from fss_b200.schemes import Dcf, Dpf

__all__ = ["Dcf", "Dpf"]
