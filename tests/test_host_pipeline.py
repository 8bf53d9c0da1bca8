"""GPU tests of the host-buffer entry points (fss_b200/csrc/host_api.cu): the adaptive pack / direct pipeline of
fssb200_eval_host, its re-entrancy (the reference's Eval is a const pure function called under
`#pragma omp parallel for`, src/bench_cpu.cu:157-161) and the arena pool.  Checker: the oracle."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np
import pytest
import torch

from oracle import HASH_KEY_BENCH, Params, synth_inputs
from test_gpu_parity import N, T, mkctx

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda:0")


def host(a):
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a)


@pytest.mark.parametrize("scheme,n,group,prg,nkeys,chunk", [
    ("dpf", 32, "bytes", "aes128_mmo", 150001, 0),       # default chunk (2^16): 3 chunks, ragged tail
    ("dpf", 32, "bytes", "aes128_mmo", 40000, 3000),     # 14 chunks: ring slots and device sets wrap
    ("halftree", 32, "u64", "aes128_mmo", 30011, 4096),
    ("dpf", 64, "u128", "chacha", 9000, 1000),
    ("dcf", 64, "u128", "aes128_mmo", 20000, 2500),      # no padding: staged copy for pageable inputs, direct for pinned
    ("dcf", 20, "bytes", "chacha", 8192, 8192),          # exactly one chunk
])
def test_pipeline_modes(dev, orc, scheme, n, group, prg, nkeys, chunk):
    p = Params(scheme=scheme, in_bits=n, group=group, prg=prg, hash_key=HASH_KEY_BENCH)
    s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=nkeys)
    o = orc.gen(p, s0s, alphas, betas, threads=8)
    oc, ooc = o if scheme == "halftree" else (o, None)
    want = orc.eval(p, 0, s0s[:, 0], oc, xs, ooc, threads=8)
    ctx = mkctx(p)
    if chunk:
        ctx.reserve_host(chunk)
    seeds_h, cws_h = host(s0s[:, 0]), host(oc)
    ocws_h = None if ooc is None else host(ooc)
    x_h = ctx.in_tensor(xs, torch.device("cpu"))
    for mode in (0, 1, 2):
        ctx.set_host_mode(mode)
        for pin_in, pin_out in ((True, True), (False, False), (True, False)):
            a = [seeds_h, cws_h, ocws_h, x_h]
            if pin_in:
                a = [None if v is None else v.pin_memory() for v in a]
            out = torch.empty((nkeys, 4), dtype=torch.int32)
            if pin_out:
                out = out.pin_memory()
            ys = ctx.eval(0, a[0], a[1], a[3], a[2], out=out)
            assert np.array_equal(N(ys), want), (mode, pin_in, pin_out)
            st = ctx.host_stats()
            assert st["packed_keys"] + st["direct_keys"] == nkeys
            if ctx.packed_row_bytes() and st["threads"] >= 2:
                if mode == 1:
                    assert st["packed_keys"] == 0
                if mode == 2 or not pin_in:
                    assert st["direct_keys"] == 0     # staged only / pageable inputs never cross as they are


def test_concurrent_host_calls(dev, orc):
    """8 threads, one shared context, host arrays: single-key calls (the `Dpf::Eval` member of the header shim) and
    large batches at the same time.  Every thread must get ITS result."""
    p = Params(scheme="dpf", in_bits=32)
    ctx = mkctx(p)
    nthreads, small, big = 8, 200, 20000
    s0s, alphas, betas, xs = synth_inputs(p, nthreads * small + 2 * big, seed=5)
    oc = orc.gen(p, s0s, alphas, betas, threads=8)
    want = orc.eval(p, 1, s0s[:, 1], oc, xs, threads=8)
    seeds_h, cws_h = host(s0s[:, 1]), host(oc)
    x_h = ctx.in_tensor(xs, torch.device("cpu"))
    got = np.zeros_like(want)
    errors = []

    def single(t):
        try:
            for i in range(t * small, (t + 1) * small):
                y = ctx.eval(1, seeds_h[i:i + 1], cws_h[i:i + 1], x_h[i:i + 1])
                got[i] = N(y)[0]
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def batch(j):
        try:
            lo = nthreads * small + j * big
            got[lo:lo + big] = N(ctx.eval(1, seeds_h[lo:lo + big], cws_h[lo:lo + big], x_h[lo:lo + big]))
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=single, args=(t,)) for t in range(nthreads)] + \
         [threading.Thread(target=batch, args=(j,)) for j in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    assert np.array_equal(got, want)


def test_arena_pool_is_sized_per_call(dev, orc):
    """A 1-key call must not reserve a batch-sized arena (round 1: ~1 GB per parameter set)."""
    from fss_b200 import _lib as L
    L.lib.fssb200_host_trim()
    d, pin = C.c_uint64(0), C.c_uint64(0)
    L.lib.fssb200_host_cached_bytes(C.byref(d), C.byref(pin))
    assert d.value == 0 and pin.value == 0
    for n, scheme in ((32, "dpf"), (128, "dcf"), (20, "halftree")):
        p = Params(scheme=scheme, in_bits=n, group="u128" if scheme == "dcf" else "bytes", hash_key=HASH_KEY_BENCH)
        ctx = mkctx(p)
        s0s, alphas, betas, xs = synth_inputs(p, 1, seed=n)
        o = orc.gen(p, s0s, alphas, betas)
        oc, ooc = o if scheme == "halftree" else (o, None)
        y = ctx.eval(0, host(s0s[:, 0]), host(oc), xs, None if ooc is None else host(ooc))
        assert np.array_equal(N(y), orc.eval(p, 0, s0s[:, 0], oc, xs, ooc))
        ctx.close()
    L.lib.fssb200_host_cached_bytes(C.byref(d), C.byref(pin))
    assert 0 < d.value <= 4 << 20 and pin.value == 0, (d.value, pin.value)
    L.lib.fssb200_host_trim()
    L.lib.fssb200_host_cached_bytes(C.byref(d), C.byref(pin))
    assert d.value == 0


def test_grotto_many_keys_scan(dev, orc):
    """More than 65535 keys through the tiled prefix-XOR scan (grid.y limit), n = 8."""
    p = Params(scheme="grotto", in_bits=8)
    ctx = mkctx(p)
    k = 70000
    s0s, alphas, _, _ = synth_inputs(p, k, seed=77)
    oc = orc.gen(p, s0s, alphas, None, threads=8)
    got = ctx.eval_all(1, T(s0s[:, 1], dev), T(oc, dev))
    want = orc.evalall(p, 1, s0s[:, 1], oc, threads=8)
    assert np.array_equal(N(got, np.uint8), want)


def test_cpp_openmp_eval(dev, tmp_path):
    """`Dpf::Eval` / `Dcf::Eval` of the header shim called from 8 OpenMP threads on one scheme object, the pattern of
    src/bench_cpu.cu:157-161 (tests/cpp/omp_eval.cpp), results against single-threaded calls and reconstruction."""
    exe = str(tmp_path / "omp_eval")
    subprocess.run(["g++", "-std=c++20", "-O1", "-fopenmp", "-I", os.path.join(ROOT, "include"), "-I",
                    "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "omp_eval.cpp"), "-o", exe, "-L",
                    os.path.join(ROOT, "fss_b200"), "-lfssb200", "-L/usr/local/cuda/lib64", "-lcudart",
                    "-Wl,-rpath," + os.path.join(ROOT, "fss_b200")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="8"))
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("prg", ["aes128_mmo", "chacha"])
@pytest.mark.parametrize("n,in_bytes,nkeys", [(1, 1, 5), (5, 1, 33), (8, 1, 1000), (10, 2, 257), (16, 2, 31), (20, 4, 4097),
                                               (32, 4, 20000), (33, 8, 100), (64, 8, 3000), (100, 16, 64), (128, 16, 999)])
def test_grotto_walk(dev, orc, prg, n, in_bytes, nkeys):
    """fssb200_grotto_eval_walk: share0 ^ share1 == 1[alpha <= x] on every key (ragged tiles, every In width, both PRGs,
    the e == 0 / e == N edge), and == the reference's Preprocess + Eval reconstructed where the parity tree exists."""
    p = Params(scheme="grotto", in_bits=n, prg=prg, in_bytes=in_bytes)
    ctx = mkctx(p)
    s0s, alphas, _, xs = synth_inputs(p, nkeys, seed=1000 + n)
    top = (1 << n) - 1
    xs[0] = top                       # e == N (or e wraps to 0 when n == 8 * in_bytes): the whole domain
    if nkeys > 4:
        xs[1], alphas[2], xs[2], alphas[3], xs[3] = 0, 0, 0, top, top
    cws = orc.gen(p, s0s, alphas, None, threads=8)
    cw_d = T(cws, dev)
    w = [N(ctx.grotto_walk(b, T(s0s[:, b], dev), cw_d, xs), np.uint8) for b in (0, 1)]
    want = np.array([1 if int(a) <= int(x) else 0 for a, x in zip(alphas, xs)], np.uint8)
    assert np.array_equal(w[0] ^ w[1], want)
    assert set(np.unique(w[0])) <= {0, 1}
    if n <= 16:
        k = min(nkeys, 8)
        ref = [orc.grotto_lookup(p, orc.grotto_preprocess(p, b, s0s[:k, b], cws[:k]), xs[:k]) for b in (0, 1)]
        assert np.array_equal((w[0] ^ w[1])[:k], ref[0] ^ ref[1])


def test_grotto_walk_rejects_other_schemes(dev):
    import fss_b200
    from fss_b200 import _lib as L
    ctx = fss_b200.Context("dpf", 16)
    buf = torch.zeros(4096, dtype=torch.int32, device=dev)
    p = C.c_void_p(buf.data_ptr())
    assert L.lib.fssb200_grotto_eval_walk(ctx.handle(0), 0, p, p, p, p, 1, None) == L.E_SCHEME
    g = fss_b200.Context("grotto", 16)
    assert L.lib.fssb200_grotto_eval_walk(g.handle(0), 2, p, p, p, p, 1, None) == L.E_INVAL
    assert L.lib.fssb200_grotto_eval_walk(g.handle(0), 0, p, p, p, p, 0, None) == 0
    with pytest.raises(RuntimeError):
        g.grotto_walk(0, torch.zeros((1, 4), dtype=torch.int32), torch.zeros((1, 17, 8), dtype=torch.int32), [0])
