"""fss_b200 -- B200-native (sm_100a) batched evaluator for the DPF / DCF / Half-Tree DPF / Grotto DCF
PRG-tree hot path of myl7/fss, behind the reference's own interfaces.

* ``Dpf`` / ``Dcf``            : drop-in for ``fss_crypto.Dpf`` / ``fss_crypto.Dcf`` (+ batched tensors)
* ``HalfTreeDpf`` / ``GrottoDcf``: the two schemes the reference binding does not expose
* ``Context``                  : thin torch wrapper of the C ABI (include/fssb200.h)

Importing this package loads fss_b200/libfssb200.so and fails loudly if it has not been built;
there is no CPU or PyTorch fallback.
"""
from . import _lib
from .context import Context, in_bytes_for, microbench
from .schemes import Dcf, Dpf, GrottoDcf, HalfTreeDpf

__all__ = ["Context", "Dcf", "Dpf", "GrottoDcf", "HalfTreeDpf", "in_bytes_for", "microbench"]
__version__ = "0.1.0"
