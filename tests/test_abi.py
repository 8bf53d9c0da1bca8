"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/fssb200.h
declares, and fails loudly (error codes, no fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fssb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fssb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fss_b200 import _lib
    names = header_symbols()
    assert len(names) >= 24
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r"\bT (fssb200_[a-z0-9_]+)", out.stdout))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(_lib.SYMBOLS), "python binding table out of sync with the header"
    hdr = open(os.path.join(ROOT, "include", "fssb200.h")).read()
    assert _lib.lib.fssb200_version() == int(re.search(r"#define FSSB200_VERSION (\d+)", hdr).group(1))


def test_library_is_sm100a_only():
    from fss_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_product_does_not_touch_the_oracle():
    """The product path must never import / link / call anything under oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "fss_b200")):
        if "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle in the GPU-less", "").replace(
                    "against the oracle", "").replace("checked against the oracle", ""), os.path.join(base, f)
    from fss_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "fssref" not in out


def test_error_codes_without_gpu():
    import torch
    from fss_b200 import _lib as L
    assert L.strerror(0) == "ok" and "scheme" in L.strerror(L.E_SCHEME)
    h = C.c_void_p()
    assert L.lib.fssb200_ctx_create(None, C.byref(h)) == L.E_INVAL
    p = L.Params()
    p.scheme, p.in_bits, p.in_bytes, p.group, p.prg = 0, 32, 4, 0, 0
    bad = L.Params.from_buffer_copy(bytes(p)); bad.scheme = 9
    assert L.lib.fssb200_ctx_create(C.byref(bad), C.byref(h)) == L.E_INVAL
    bad = L.Params.from_buffer_copy(bytes(p)); bad.in_bits = 33
    assert L.lib.fssb200_ctx_create(C.byref(bad), C.byref(h)) == L.E_DOMAIN
    bad = L.Params.from_buffer_copy(bytes(p)); bad.in_bytes = 3
    assert L.lib.fssb200_ctx_create(C.byref(bad), C.byref(h)) == L.E_INVAL
    bad = L.Params.from_buffer_copy(bytes(p)); bad.group = 5  # U128 needs 0 < mod <= 2^127
    assert L.lib.fssb200_ctx_create(C.byref(bad), C.byref(h)) == L.E_GROUP
    bad = L.Params.from_buffer_copy(bytes(p)); bad.group = 0; bad.mod_lo = 7  # Bytes has no modulus
    assert L.lib.fssb200_ctx_create(C.byref(bad), C.byref(h)) == L.E_GROUP
    if not torch.cuda.is_available():
        # no device: loud failure, never a CPU fallback
        assert L.lib.fssb200_ctx_create(C.byref(p), C.byref(h)) == L.E_NODEVICE
        v = C.c_double()
        assert L.lib.fssb200_microbench(0, 0, C.byref(v)) == L.E_NODEVICE
    assert L.lib.fssb200_ctx_ncw(None) == L.E_INVAL
    assert L.lib.fssb200_eval(None, 0, None, None, None, None, None, 0, None) == L.E_INVAL


@pytest.mark.parametrize("src,extra", [("shim_sample.cpp", []), ("omp_eval.cpp", ["-fopenmp"])])
def test_header_shim_compiles_with_plain_gxx(tmp_path, src, extra):
    """include/fss/*.cuh with the built-in plugins needs no nvcc: g++ -std=c++20 + the C-ABI library (compile and link
    only here; the -m gpu tests run the binaries)."""
    exe = str(tmp_path / src.replace(".cpp", ""))
    subprocess.run(["g++", "-std=c++20", "-O1", *extra, "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "cpp", src), "-o", exe, "-L", os.path.join(ROOT, "fss_b200"), "-lfssb200",
                    "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "fss_b200")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    import torch
    if not torch.cuda.is_available():   # fails loudly without a GPU: an exception from fssb200_ctx_create, never a CPU result
        assert r.returncode != 0 and "no such CUDA device" in (r.stdout + r.stderr)


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md maps every exported function of include/fssb200.h to the reference interface it replaces."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    syms = set(re.findall(r"\b(fssb200_[a-z0-9_]+)\(", open(os.path.join(root, "include", "fssb200.h")).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = sorted(s for s in syms if s not in doc)
    assert not missing, missing
