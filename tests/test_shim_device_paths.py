"""The header shim's DEVICE-callable members beyond DPF / DCF / Half-Tree with ChaCha: `prg::Aes128Soft` on the caller's
tables, `hash::Blake3`, `Vdpf::Gen` / `Vdpf::Eval` (include/fss/prg/aes128_mmo_soft.cuh, hash/blake3.cuh, vdpf.cuh,
b200/generic.cuh) -- what the reference's own GPU benchmark (src/bench_gpu.cu) calls per thread inside its kernels.

* tests/host_emul/shim_device_paths.cpp: the `__host__ __device__` functions those members run on the device, compiled for the
  host and compared bit for bit with the oracle and the survey's known answers (CPU);
* the reference's src/bench_gpu.cu, unmodified, cross-compiled for sm_100a against include/ of this repository by
  oracle/Makefile (oracle/gbench_stub/ stands in for the Google Benchmark headers the reference fetches from the network).
* tests/cpp/device_members.cu (-m gpu): kernels calling those members per thread, bit-compared with the batched members behind
  the C ABI on the B200 (passed on hardware with the round's last GPU seconds: profiles/r02_last_session_gpu_checks.md).
The reference's benchmark binary itself (oracle/_ref/reftests/bench_gpu) has been built, not timed on a B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_device_path_functions_match_the_oracle_on_the_host(tmp_path):
    exe = str(tmp_path / "shim_device_paths")
    subprocess.run(["g++", "-std=c++20", "-O1", "-w", "-x", "c++", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "host_emul", "shim_device_paths.cpp"), "-o", exe, "-L", os.path.join(ROOT, "oracle"),
                    "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "shim device paths: all checks passed" in r.stdout and "FAIL" not in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


def test_reference_gpu_benchmark_is_built_unmodified_for_sm100a():
    """oracle/Makefile (`reftests`, part of build()) compiles the reference's src/bench_gpu.cu from where it lies against include/;
    its kernels -- the per-thread Gen / Eval with Aes128Soft, VDPF + Blake3, ChaCha -- are in the binary as sm_100a code."""
    exe = os.path.join(ROOT, "oracle", "_ref", "reftests", "bench_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/reftests/bench_gpu not built (needs the reference checkout: make -C oracle reftests)")
    r = subprocess.run(["cuobjdump", "-lelf", exe], capture_output=True, text=True)
    assert "sm_100a" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    names = subprocess.run(["cuobjdump", "-elf", exe], capture_output=True, text=True).stdout
    for kernel in ("DpfEvalKernelAes", "DpfGenKernelAes", "VdpfGenKernel", "VdpfEvalKernel", "DcfEvalKernel", "HalfTreeDpfEvalKernel"):
        assert kernel in names, kernel


def test_reference_benchmark_programs_are_built():
    """oracle/Makefile builds the reference's three benchmark programs, unmodified, against include/ + libfssb200.so."""
    if not os.path.isdir(os.path.join(REF, "include", "fss")):
        pytest.skip("reference checkout not present")
    for name in ("bench_gpu", "bench_cpu", "bench_xlib"):
        assert os.path.getsize(os.path.join(ROOT, "oracle", "_ref", "reftests", name)) > 0, name


def build_device_members(out):
    subprocess.run(["nvcc", "-std=c++20", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "device_members.cu"), "-o", out, "-L", os.path.join(ROOT, "fss_b200"), "-lfssb200",
                    "-Xlinker", "-rpath," + os.path.join(ROOT, "fss_b200")], check=True)


def test_device_members_program_compiles_for_sm100a(tmp_path):
    """tests/cpp/device_members.cu: kernels that call Dpf<Aes128Soft>::Gen / Eval (tables in shared memory) and Vdpf::Gen / Eval
    (ChaCha + Blake3) per thread, compared with the batched members behind the C ABI."""
    exe = str(tmp_path / "device_members")
    build_device_members(exe)
    assert os.path.getsize(exe) > 0


@pytest.mark.gpu
def test_device_members_on_the_gpu(tmp_path):
    exe = str(tmp_path / "device_members")
    build_device_members(exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "device members: all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
