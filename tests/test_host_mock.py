"""CPU test of the product's host-side pipeline (fss_b200/csrc/host_api.cu, unmodified) against a mock CUDA runtime:
worker crew, staging ring, piece / chunk hand-offs, adaptive direct pieces, arena pool, multi-device host call (equal and
balanced split), the chunked loops of gen / prg_gen / eval_all / eval_levelmajor / VDPF host calls, error paths -- with
"kernels" that digest every byte of every key, so a result is right only if the data reached the mock device intact and in
the format the launch claimed.  The same binary also runs under ThreadSanitizer and under AddressSanitizer + UBSan (the mock
device memory is plain heap memory: an offset error inside a device set is an ASAN report).
(The real kernels behind the same entry points are covered by the -m gpu tests.)"""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "host_pipe_mock.cpp")
INC = ["-I", os.path.join(HERE, "host_emul"), "-I", os.path.join(os.path.dirname(HERE), "fss_b200", "csrc")]


def build(out, extra):
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", *extra, "-x", "c++", *INC, SRC, "-o", out], check=True)


@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_host_pipeline_against_mock_runtime(tmp_path, threads):
    exe = str(tmp_path / "host_pipe_mock")
    build(exe, [])
    r = subprocess.run([exe] + (["quick"] if threads != "3" else []), capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, FSSB200_PACK_THREADS=threads))
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_host_pipeline_is_tsan_clean(tmp_path):
    exe = str(tmp_path / "host_pipe_mock_tsan")
    build(exe, ["-fsanitize=thread"])
    r = subprocess.run([exe, "quick"], capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, FSSB200_PACK_THREADS="6", TSAN_OPTIONS="halt_on_error=0 exitcode=66"))
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-6000:]
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_host_pipeline_is_asan_and_ubsan_clean(tmp_path):
    exe = str(tmp_path / "host_pipe_mock_asan")
    build(exe, ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"])
    r = subprocess.run([exe, "quick"], capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, FSSB200_PACK_THREADS="6", ASAN_OPTIONS="detect_leaks=0 exitcode=67"))
    assert "AddressSanitizer" not in r.stderr, r.stderr[-6000:]
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
