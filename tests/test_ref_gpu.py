"""GPU-side second oracle: the reference's OWN device kernels (point_eval_gpu.cuh, eval_all_gpu.cuh, the
Aes128Soft / ChaCha device PRGs), compiled unmodified for sm_100a into oracle/_ref/ref_gpu_bench by
`make -C oracle refgpu`, run on the same keys as libfssb200.so.  Bar: bit-exact.

The reference ships no test for its point-eval kernels and one ad-hoc check for EvalAll
(check_evalall_gpu.cu); this file is that check, restated against this library.
"""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench")

POINT = [  # mode, scheme, in_bits, group, prg, wide xs
    ("dpf32_chacha_naive", "dpf", 32, "bytes", "chacha", False),
    ("dpf32_chacha_point", "dpf", 32, "bytes", "chacha", False),
    ("dpf32_aes_naive", "dpf", 32, "bytes", "aes128_mmo", False),
    ("dcf64_u127_aes_naive", "dcf", 64, "u128", "aes128_mmo", True),
    ("dcf64_u127_chacha_naive", "dcf", 64, "u128", "chacha", True),
    ("dcf32_u64_chacha_point", "dcf", 32, "u64", "chacha", False),
    ("ht32_chacha_point", "halftree", 32, "bytes", "chacha", False),
    ("ht32_aes_naive", "halftree", 32, "bytes", "aes128_mmo", False),
]
EVALALL = [("dpf_evalall20_chacha", "dpf", 20, 3), ("ht_evalall20_chacha", "halftree", 20, 3)]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ref_gpu_bench not built (needs /root/reference at build time)")
    return torch.device("cuda:0")


def _rand(shape, gen, dev, wide=False):
    if wide:
        return torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device=dev, generator=gen)
    return torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int64, device=dev, generator=gen).to(torch.int32)


def _run(mode, nkeys, d):
    r = subprocess.run([BIN, mode, str(nkeys), str(d), "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode,scheme,n,group,prg,wide", POINT, ids=[p[0] for p in POINT])
def test_point_eval_matches_reference_gpu_kernels(dev, tmp_path, mode, scheme, n, group, prg, wide):
    import fss_b200
    nkeys = 5000  # ragged: not a multiple of the reference's 256-thread blocks nor of this library's 32-key tiles
    gen = torch.Generator(device=dev).manual_seed(99)
    ctx = fss_b200.Context(scheme, n, group, prg=prg)
    s0s, betas = _rand((nkeys, 2, 4), gen, dev), _rand((nkeys, 4), gen, dev)
    s0s[:, :, 3] &= ~1
    betas[:, 3] &= ~1
    alphas, xs = _rand((nkeys,), gen, dev, wide), _rand((nkeys,), gen, dev, wide)
    xs[::7] = alphas[::7]
    r = ctx.gen(s0s, alphas, betas)
    cws, ocws = r if scheme == "halftree" else (r, None)
    seeds0 = s0s[:, 0].contiguous()
    for name, t in (("seeds", seeds0), ("cws", cws), ("xs", xs), ("ocws", ocws)):
        if t is not None:
            t.contiguous().cpu().numpy().tofile(tmp_path / f"{name}.bin")
    _run(mode, nkeys, tmp_path)
    want = np.fromfile(tmp_path / "ys_ref.bin", dtype=np.int32).reshape(nkeys, 4)
    got = ctx.eval(0, seeds0, cws, xs, ocws=ocws).cpu().numpy()
    assert np.array_equal(got, want)
    if mode.endswith("_point"):
        got2 = ctx.eval_levelmajor(0, seeds0, ctx.relayout(cws), xs, ocws=ocws).cpu().numpy()
        assert np.array_equal(got2, want)


@pytest.mark.parametrize("mode,scheme,n,nkeys", EVALALL, ids=[e[0] for e in EVALALL])
def test_eval_all_matches_reference_gpu_kernels(dev, tmp_path, mode, scheme, n, nkeys):
    import fss_b200
    gen = torch.Generator(device=dev).manual_seed(7)
    ctx = fss_b200.Context(scheme, n, "bytes", prg="chacha")
    s0s, betas = _rand((nkeys, 2, 4), gen, dev), _rand((nkeys, 4), gen, dev)
    s0s[:, :, 3] &= ~1
    betas[:, 3] &= ~1
    alphas = _rand((nkeys,), gen, dev) & ((1 << n) - 1)
    r = ctx.gen(s0s, alphas, betas)
    cws, ocws = r if scheme == "halftree" else (r, None)
    seeds0 = s0s[:, 0].contiguous()
    for name, t in (("seeds", seeds0), ("cws", cws), ("ocws", ocws)):
        if t is not None:
            t.contiguous().cpu().numpy().tofile(tmp_path / f"{name}.bin")
    _run(mode, nkeys, tmp_path)
    want = np.fromfile(tmp_path / "ys_ref.bin", dtype=np.int32).reshape(nkeys, 1 << n, 4)
    got = ctx.eval_all(0, seeds0, cws, ocws=ocws).cpu().numpy()
    assert np.array_equal(got, want)
