"""ctypes binding of fss_b200/libfssb200.so (the C ABI declared in include/fssb200.h).

The library is the only evaluation path: if it is missing, import fails loudly -- there is no
CPU or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfssb200.so")

SCHEME_DPF, SCHEME_DCF, SCHEME_HALFTREE, SCHEME_GROTTO, SCHEME_VDPF = 0, 1, 2, 3, 4
GROUP_BYTES, GROUP_U8, GROUP_U16, GROUP_U32, GROUP_U64, GROUP_U128 = 0, 1, 2, 3, 4, 5
PRG_AES128_MMO, PRG_CHACHA = 0, 1
PRED_LT, PRED_GT = 0, 1

E_INVAL, E_DOMAIN, E_GROUP, E_SCHEME, E_ALIGN, E_NODEVICE, E_RANGE, E_NOARENA = -1, -2, -3, -4, -5, -6, -7, -8


class Params(C.Structure):
    """``fssb200_params`` (include/fssb200.h)."""
    _fields_ = [("scheme", C.c_int32), ("in_bits", C.c_int32), ("in_bytes", C.c_int32), ("group", C.c_int32),
                ("mod_lo", C.c_uint64), ("mod_hi", C.c_uint64), ("prg", C.c_int32), ("pred", C.c_int32),
                ("prg_key", C.c_uint8 * 64), ("hash_key", C.c_uint8 * 16), ("device", C.c_int32),
                ("hash", C.c_int32), ("hash_iv", C.c_uint8 * 64)]


class FssError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what} failed: {strerror(code)} (code {code})")


# every symbol include/fssb200.h declares: (name, restype, argtypes)
_VP, _SZ, _U64, _I = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
SYMBOLS = {
    "fssb200_version": (_I, []),
    "fssb200_strerror": (C.c_char_p, [_I]),
    "fssb200_ctx_create": (_I, [C.POINTER(Params), C.POINTER(_VP)]),
    "fssb200_ctx_destroy": (None, [_VP]),
    "fssb200_ctx_params": (_I, [_VP, C.POINTER(Params)]),
    "fssb200_ctx_ncw": (_I, [_VP]),
    "fssb200_gen": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_eval": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_dpf_eval": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_dcf_eval": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_halftree_eval": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_eval_all": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _SZ, _U64, _U64, _VP]),
    "fssb200_eval_all_granule": (_U64, [_VP]),
    "fssb200_grotto_expand": (_I, [_VP, _I, _VP, _VP, _VP, _SZ, _U64, _U64, _VP]),
    "fssb200_grotto_preprocess": (_I, [_VP, _I, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_grotto_eval": (_I, [_VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_grotto_eval_walk": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_vdpf_gen": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_vdpf_eval": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_vdpf_eval_levelmajor": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_vdpf_prove": (_I, [_VP, _VP, _VP, _SZ, _VP, _SZ, _VP]),
    "fssb200_vdpf_eval_all": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_hash": (_I, [_VP, _I, _VP, _VP, _SZ, _VP]),
    "fssb200_vdpf_gen_host": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ]),
    "fssb200_vdpf_eval_host": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ]),
    "fssb200_relayout": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_eval_levelmajor": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_ctx_reserve_host": (_I, [_VP, _SZ]),
    "fssb200_ctx_set_host_mode": (_I, [_VP, _I]),
    "fssb200_ctx_host_stats": (_I, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "fssb200_host_trim": (None, []),
    "fssb200_host_cached_bytes": (_I, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fssb200_eval_host": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _SZ]),
    "fssb200_eval_all_host": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _SZ, _U64, _U64]),
    "fssb200_gen_host": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ]),
    "fssb200_eval_levelmajor_host": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ]),
    "fssb200_eval_multi": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "fssb200_eval_all_multi": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "fssb200_gen_multi": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "fssb200_multi_sync": (_I, [_VP, _I, _VP, _VP]),
    "fssb200_key_shard": (_I, [_SZ, _I, _I, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "fssb200_leaf_shard": (_I, [_VP, _I, _I, C.POINTER(_U64), C.POINTER(_U64)]),
    "fssb200_eval_host_multi": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_packed_row_bytes": (_SZ, [_VP]),
    "fssb200_ctx_host_pack_threads": (_I, [_VP]),
    "fssb200_pack_rows": (_I, [_VP, _VP, _VP, _SZ]),
    "fssb200_eval_packed": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "fssb200_prg_gen": (_I, [_VP, _VP, _VP, _I, _SZ, _VP]),
    "fssb200_prg_gen_host": (_I, [_VP, _VP, _VP, _I, _SZ]),
    "fssb200_ctx_launch_count": (_U64, [_VP]),
    "fssb200_microbench": (_I, [_I, _I, C.POINTER(C.c_double)]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a extension first "
            "(`python -c 'import __graft_entry__ as g; g.build()'` or `make -C fss_b200/csrc -j8`). "
            "fss_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def strerror(code: int) -> str:
    return lib.fssb200_strerror(code).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise FssError(code, what)
