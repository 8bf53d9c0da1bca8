"""The reference's OWN googletest suites (src/{dpf,dcf,half_tree_dpf,grotto_dcf,vdpf,vdmpf,group}_test.cu), compiled UNMODIFIED --
but against this repository's header tree and linked with libfssb200.so (oracle/Makefile: `make reftests`, built in the
container where the reference checkout is; the binaries travel to the GPU box under oracle/_ref/reftests/).  A user of the
reference switches by changing one include path and one library; this is that switch applied to the reference's own
tests: 33 group-axiom tests on the CPU, 65 scheme tests (11 + 7 + 27 + 5 + 8 + 7: reconstruction at / off alpha, EvalAll, Grotto edge cases, VDPF / VDMPF
verification; ChaCha, Aes128Mmo and Aes128Soft PRGs; Bytes / Uint64 / Uint127 groups; in_bits = 1) on the B200."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "reftests")


def run(name):
    exe = os.path.join(BIN, name + "_test")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs the reference checkout: make -C oracle reftests)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    m = re.search(r"\[  PASSED  \] (\d+) tests", r.stdout)
    assert r.returncode == 0 and m and "FAILED" not in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
    return int(m.group(1))


def test_reference_group_axioms_on_the_shim_groups():
    assert run("group") >= 30


GOMP = ["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"]


def link_over_the_oracle_backend(tmp_path, obj, exe, extra=()):
    """Links a test object of the reference (built against include/ of this repository by oracle/Makefile) with
    tests/host_emul/fake_backend.cpp: the C ABI answered by the oracle on the CPU instead of libfssb200.so."""
    fake = str(tmp_path / "fake_backend.o")
    subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I", "/usr/local/cuda/include", "-c",
                    os.path.join(ROOT, "tests", "host_emul", "fake_backend.cpp"), "-o", fake], check=True)
    subprocess.run(["g++", obj, fake, *extra, "-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                    *GOMP, "-lpthread", "-o", exe], check=True)


@pytest.mark.parametrize("name,tests", [("dpf", 11), ("dcf", 7), ("half_tree_dpf", 27), ("grotto_dcf", 5), ("vdpf", 8), ("vdmpf", 7)])
def test_reference_gtest_suite_on_the_cpu_over_the_oracle_backend(tmp_path, name, tests):
    """The header shim's host side without a GPU: the reference's suites, unmodified, with every C-ABI batch answered by the
    oracle.  What this pins is the shim -- parameter marshalling for every scheme x group x PRG the suites instantiate, key
    layouts, the multi-point host logic; the kernels behind the same calls are pinned by the -m gpu run of the same sources."""
    obj = os.path.join(BIN, name + "_test.o")
    gtest = os.path.join(ROOT, "oracle", "_ref", "gtest", "lib")
    if not os.path.exists(obj) or not os.path.exists(os.path.join(gtest, "libgtest.a")):
        pytest.skip(f"{obj} not built (needs the reference checkout: make -C oracle reftests)")
    exe = str(tmp_path / (name + "_test_cpu"))
    link_over_the_oracle_backend(tmp_path, obj, exe, [os.path.join(gtest, "libgtest_main.a"), os.path.join(gtest, "libgtest.a")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"[  PASSED  ] {tests} tests" in r.stdout and "FAILED" not in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["dpf_dcf_cpu", "half_tree_dpf_cpu", "grotto_dcf_cpu", "vdpf_cpu", "vdmpf_cpu"])
def test_reference_sample_on_the_cpu_over_the_oracle_backend(tmp_path, name):
    obj = os.path.join(BIN, "sample_" + name + ".o")
    if not os.path.exists(obj):
        pytest.skip(f"{obj} not built (needs the reference checkout: make -C oracle reftests)")
    exe = str(tmp_path / ("sample_" + name + "_cpu"))
    link_over_the_oracle_backend(tmp_path, obj, exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    out = r.stdout
    assert r.returncode == 0 and "===" in out, out[-3000:] + r.stderr[-2000:]
    assert "? no" not in out and not re.search(r"\?\s+NO\b", out) and not re.search(r"mismatches[^\n]*: [1-9]", out), out


@pytest.mark.gpu
@pytest.mark.parametrize("name,at_least", [("dpf", 10), ("dcf", 7), ("half_tree_dpf", 27), ("grotto_dcf", 5), ("vdpf", 8), ("vdmpf", 7)])
def test_reference_gtest_suite_passes_against_this_library(name, at_least):
    assert run(name) >= at_least


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dpf_dcf_cpu", "half_tree_dpf_cpu", "grotto_dcf_cpu", "vdpf_cpu", "vdmpf_cpu", "dpf_dcf_gpu"])
def test_reference_sample_runs_against_this_library(name):
    """The reference's samples/*.cu, unmodified, built against include/ of this repository.  dpf_dcf_gpu.cu is the
    documented in-kernel usage (README.md:198-242): `dpf.Gen` / `dpf.Eval` called per thread inside the sample's own
    __global__ kernels with the ChaCha PRG.  vdpf_cpu.cu uses fss::hash::Sha256 for both hashes (host-only in the reference,
    on the device here)."""
    exe = os.path.join(BIN, "sample_" + name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs the reference checkout: make -C oracle reftests)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    out = r.stdout
    assert r.returncode == 0 and "===" in out, out[-3000:] + r.stderr[-2000:]
    assert "? no" not in out and not re.search(r"\?\s+NO\b", out) and not re.search(r"mismatches[^\n]*: [1-9]", out), out
    for m in re.finditer(r"Verification: (\d+)/(\d+) correct", out):
        assert m.group(1) == m.group(2), out


@pytest.mark.gpu
def test_reference_cpu_gpu_parity_check_runs_against_this_library():
    """check_evalall_gpu.cu (the only CPU == GPU check in the reference's tree), unmodified:
    the free functions fss::gpu::DpfEvalAllGpu / HalfTreeDpfEvalAllGpu on device arrays against the members' EvalAll on host
    arrays (n = 20, Uint<uint64_t>, ChaCha), and the reconstruction of the device outputs at / off alpha."""
    exe = os.path.join(BIN, "check_evalall_gpu")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs the reference checkout: make -C oracle reftests)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL PASS" in r.stdout and "FAIL" not in r.stdout.replace("FAILURES", ""), r.stdout[-3000:] + r.stderr[-2000:]
    assert r.stdout.count("PASS") >= 6, r.stdout


def test_reference_cpu_benchmark_on_the_cpu_over_the_oracle_backend(tmp_path):
    """The reference's src/bench_cpu.cu (Google Benchmark harness: oracle/gbench_stub stands in for the library the reference
    fetches from the network), unmodified, on the shim headers: built by oracle/Makefile, linked here with the oracle-backed
    C ABI, a subset of its 32 benchmarks run once -- every scheme's Gen, and the OpenMP loop over keys calling Dpf::Eval."""
    obj, main_o = os.path.join(BIN, "bench_cpu.o"), os.path.join(BIN, "gbench_main.o")
    if not os.path.exists(obj) or not os.path.exists(main_o):
        pytest.skip(f"{obj} not built (needs the reference checkout: make -C oracle reftests)")
    exe = str(tmp_path / "bench_cpu_cpu")
    link_over_the_oracle_backend(tmp_path, obj, exe, [main_o])
    for flt, at_least in (("Gen", 8), ("BM_DpfEval_Uint_Aes/14", 1)):
        r = subprocess.run([exe, flt], capture_output=True, text=True, timeout=600, env=dict(os.environ, FSS_BENCH_ITERS="1"))
        lines = [l for l in r.stdout.splitlines() if l.startswith("BM_")]
        assert r.returncode == 0 and len(lines) >= at_least and all("(1 iterations)" in l for l in lines), r.stdout[-3000:] + r.stderr[-2000:]
