"""Plugin extensibility and device-callable members of the header shim (include/fss/*.cuh, fss/b200/generic.cuh).

tests/cpp/plugin_user.cu instantiates the reference's class templates with a USER-DEFINED group and PRG
(tests/cpp/user_plugin.hpp: they satisfy group.cuh:39-45 / prg.cuh:20-23 and nothing else) and calls `Gen` / `Eval`
both from the host and from inside its own __global__ kernels (README.md:198-242).  The golden file is the output of the
SAME source compiled against the reference's unmodified headers on the CPU (oracle/make_golden_plugin.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "plugin_user_v1.txt")


def build(out):
    subprocess.run(["nvcc", "-std=c++20", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "plugin_user.cu"), "-o", out, "-L", os.path.join(ROOT, "fss_b200"),
                    "-lfssb200", "-Xlinker", "-rpath," + os.path.join(ROOT, "fss_b200")], check=True)


def test_user_plugin_translation_unit_compiles_for_sm100a(tmp_path):
    """nvcc cross-compiles the user's translation unit (generic kernels instantiated with the user's types) without a GPU."""
    exe = str(tmp_path / "plugin_user")
    build(exe)
    assert os.path.getsize(exe) > 0


def test_golden_is_the_reference_output():
    """Where the reference checkout exists (build container), the committed golden file is what the reference prints."""
    if not os.path.isdir("/root/reference/include/fss"):
        pytest.skip("reference checkout not present")
    before = open(GOLDEN).read()
    subprocess.run(["python", os.path.join(ROOT, "oracle", "make_golden_plugin.py")], check=True, capture_output=True)
    assert open(GOLDEN).read() == before
    assert "BAD" not in before and before.count("reconstruct") == before.count(" ok")


@pytest.mark.gpu
def test_user_plugins_and_device_members_on_the_gpu(tmp_path):
    exe = str(tmp_path / "plugin_user")
    build(exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "plugin test: all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    want = open(GOLDEN).read().splitlines()
    got = [l for l in r.stdout.splitlines() if not l.startswith(("shim:", "plugin test:"))]
    assert got == want, next((i, a, b) for i, (a, b) in enumerate(zip(got + [""], want + [""])) if a != b)
