"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through the C ABI
(fss_b200.Context -> libfssb200.so); the oracle and the reference-generated golden fixtures are the
checkers.  Bar: bit-exact.
"""
import ctypes as C
import hashlib

import numpy as np
import pytest
import torch

from oracle import HASH_KEY_BENCH, Params, synth_inputs

pytestmark = pytest.mark.gpu

GROUP_NAME = {"bytes": "bytes", "u8": "u8", "u16": "u16", "u32": "u32", "u64": "u64", "u128": "u128"}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda:0")


def mkctx(p: Params):
    import fss_b200
    return fss_b200.Context(p.scheme, p.in_bits, GROUP_NAME[p.group], p.mod, p.prg, p.pred, p.prg_key, p.hash_key,
                            p.in_bytes)


def T(a, dev):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(dev)


def N(t, dtype=np.uint32):
    return t.detach().cpu().contiguous().numpy().view(dtype)


def masked(p, cws):
    from golden_util import Case
    c = Case.__new__(Case)
    c.p = p
    return c.masked_cws(np.ascontiguousarray(cws))


# ---- known answers -----------------------------------------------------------------------------------------

def test_prg_known_answers(dev, golden, orc):
    seeds = golden.arrays["prg/seeds"]
    for prg in ("aes128_mmo", "chacha"):
        ctx = mkctx(Params(scheme="dcf", in_bits=8, prg=prg))
        for mul in (1, 2, 4):
            got = N(ctx.prg_gen(T(seeds, dev), mul))
            assert np.array_equal(got, golden.arrays[f"prg/{prg}_{mul}"]), (prg, mul)
    # a batch large enough to use every SM, random seeds
    big = np.random.default_rng(3).integers(0, 2 ** 32, size=(1 << 16, 4), dtype=np.uint64).astype(np.uint32)
    for prg in ("aes128_mmo", "chacha"):
        p = Params(scheme="dcf", in_bits=8, prg=prg)
        assert np.array_equal(N(mkctx(p).prg_gen(T(big, dev), 4)), orc.prg_gen(p, 4, big)), prg


def test_golden_cases(dev, golden):
    """Every reference-generated fixture: Gen, Eval (both parties), EvalAll / Grotto."""
    ran = 0
    for c in golden.cases:
        p = c.p
        ctx = mkctx(p)
        s0s = T(c["s0s"], dev)
        r = ctx.gen(s0s, c.alphas, None if p.scheme == "grotto" else T(c["betas"], dev))
        cws, ocws = r if p.scheme == "halftree" else (r, None)
        assert np.array_equal(masked(p, N(cws)), c.masked_cws(c["cws"])), ("gen", c.name)
        if ocws is not None:
            assert np.array_equal(N(ocws), c["ocws"]), ("ocw", c.name)
        gcws = T(c["cws"], dev)  # evaluate the REFERENCE's keys
        gocws = T(c["ocws"], dev) if p.scheme == "halftree" else None
        for party in (0, 1):
            seeds = s0s[:, party].contiguous()
            if p.scheme != "grotto":
                ys = ctx.eval(party, seeds, gcws, c.xs, gocws)
                assert np.array_equal(N(ys), c[f"ys{party}"]), ("eval", c.name, party)
            mode = c.meta["evalall"]
            if mode == "none" or p.in_bits > 24:
                continue  # n = 28 has its own test below
            k = c.meta["evalall_keys"]
            ya = ctx.eval_all(party, seeds[:k], gcws[:k], None if gocws is None else gocws[:k])
            ya = N(ya, np.uint8 if p.scheme == "grotto" else np.uint32)
            if mode == "full":
                assert np.array_equal(ya, c[f"all{party}"]), ("evalall", c.name, party)
            else:
                digests = [hashlib.sha256(np.ascontiguousarray(ya[i]).tobytes()).hexdigest() for i in range(k)]
                assert digests == c.meta[f"all{party}_sha256"], ("evalall sha", c.name, party)
            if c.has(f"pt{party}"):
                pt = ctx.grotto_preprocess(party, seeds[:k], gcws[:k])
                assert np.array_equal(N(pt, np.uint8), c[f"pt{party}"]), ("pt", c.name)
                assert np.array_equal(N(ctx.grotto_lookup(pt, c.xs[:k]), np.uint8), c[f"lookup{party}"]), c.name
        ran += 1
    assert ran >= 100


# ---- random batches against the oracle --------------------------------------------------------------------------

CONFIGS = [
    # scheme, n, group, mod, prg, nkeys
    ("dpf", 32, "bytes", 0, "aes128_mmo", 1 << 15),            # C2
    ("dcf", 64, "u128", 1 << 127, "aes128_mmo", 1 << 13),      # C3
    ("halftree", 32, "bytes", 0, "aes128_mmo", 1 << 14),       # C5
    ("dpf", 32, "bytes", 0, "chacha", 1 << 14),
    ("dcf", 64, "u128", 1 << 127, "chacha", 1 << 12),
    ("halftree", 32, "u64", 0, "chacha", 1 << 13),
    ("dpf", 20, "u64", 0, "aes128_mmo", 5000),                  # ragged batch (not a multiple of 32 / 512)
    ("dcf", 32, "u32", 4294967291, "aes128_mmo", 3001),
    ("dcf", 16, "u8", 251, "chacha", 777),
    ("dpf", 128, "u128", (1 << 127) - 1, "aes128_mmo", 1000),
    ("dcf", 100, "u64", 18446744073709551557, "aes128_mmo", 1000),
    ("halftree", 1, "u64", 0, "aes128_mmo", 100),
    ("dpf", 1, "bytes", 0, "chacha", 33),
    ("dcf", 33, "u16", 0, "aes128_mmo", 513),
]


@pytest.mark.parametrize("scheme,n,group,mod,prg,nkeys", CONFIGS)
def test_random_batches(dev, orc, scheme, n, group, mod, prg, nkeys):
    p = Params(scheme=scheme, in_bits=n, group=group, mod=mod, prg=prg, hash_key=HASH_KEY_BENCH)
    ctx = mkctx(p)
    s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=n * 31 + nkeys)
    xs[1], xs[2] = 0, (1 << n) - 1
    r = ctx.gen(T(s0s, dev), alphas, T(betas, dev))
    cws, ocws = r if scheme == "halftree" else (r, None)
    o = orc.gen(p, s0s, alphas, betas, threads=8)
    oc, ooc = o if scheme == "halftree" else (o, None)
    assert np.array_equal(N(cws), oc)
    if ocws is not None:
        assert np.array_equal(N(ocws), ooc)
    for party in (0, 1):
        ys = ctx.eval(party, T(s0s[:, party], dev), cws, xs, ocws)
        assert np.array_equal(N(ys), orc.eval(p, party, s0s[:, party], oc, xs, ooc, threads=8)), party
    # level-major pre-pass (fss::gpu::*RelayoutGpu + *EvalPointGpu semantics)
    lay = ctx.relayout(cws)
    want = orc.relayout(p, oc)
    assert np.array_equal(N(lay[0]), want[0])
    if scheme == "dcf":
        assert np.array_equal(N(lay[1]), want[1])
    else:
        assert np.array_equal(N(lay[2]), want[2])
    if scheme != "halftree":
        assert np.array_equal(N(lay[3]), want[3])
    ys_lm = ctx.eval_levelmajor(1, T(s0s[:, 1], dev), lay, xs, ocws)
    want_ys = N(ctx.eval(1, T(s0s[:, 1], dev), cws, xs, ocws))
    assert np.array_equal(N(ys_lm), want_ys)
    # the same layout from HOST arrays (fssb200_eval_levelmajor_host), in chunks smaller than the batch
    ctx.reserve_host(max(1, nkeys // 3))
    lay_h = tuple(None if t is None else t.cpu() for t in lay)
    ys_h = ctx.eval_levelmajor(1, T(s0s[:, 1], "cpu"), lay_h, xs, None if ocws is None else ocws.cpu())
    assert ys_h.device.type == "cpu" and np.array_equal(N(ys_h), want_ys)
    assert ctx.launch_count() >= 5


@pytest.mark.parametrize("scheme,n,group,prg,nkeys", [
    ("dpf", 18, "bytes", "aes128_mmo", 3), ("dpf", 10, "u64", "aes128_mmo", 9), ("dpf", 5, "u128", "chacha", 70),
    ("halftree", 18, "u64", "aes128_mmo", 2), ("halftree", 9, "bytes", "chacha", 5), ("halftree", 1, "u32", "aes128_mmo", 4),
    ("grotto", 18, "bytes", "aes128_mmo", 2), ("grotto", 11, "bytes", "chacha", 5), ("dpf", 1, "bytes", "aes128_mmo", 3),
    # Grotto, bit-packed leaf output (n >= 14) into rows of every alignment (parity trees are 2N-1 bytes apart)
    ("grotto", 14, "bytes", "aes128_mmo", 5), ("grotto", 15, "bytes", "chacha", 3), ("grotto", 19, "bytes", "aes128_mmo", 3),
    ("grotto", 13, "bytes", "aes128_mmo", 4), ("grotto", 1, "bytes", "aes128_mmo", 2), ("grotto", 20, "bytes", "chacha", 2),
    ("dpf", 19, "u128", "chacha", 2), ("dcf", 17, "u128", "aes128_mmo", 2), ("dcf", 10, "u64", "chacha", 5),
    ("dcf", 3, "bytes", "aes128_mmo", 9), ("dcf", 12, "u32", "aes128_mmo", 3),
    # cooperative bottom stage of the leaf kernels (n >= 12): shortest walks (dfs = 3, 4, 5), both PRGs, every leaf mode
    ("dpf", 12, "u32", "aes128_mmo", 5), ("halftree", 12, "u64", "chacha", 3), ("halftree", 13, "bytes", "aes128_mmo", 4),
    ("dpf", 14, "u128", "chacha", 2), ("dpf", 13, "bytes", "aes128_mmo", 150), ("dpf", 11, "u64", "aes128_mmo", 7),
])
def test_evalall_vs_oracle(dev, orc, scheme, n, group, prg, nkeys):
    p = Params(scheme=scheme, in_bits=n, group=group, prg=prg, hash_key=HASH_KEY_BENCH)
    ctx = mkctx(p)
    s0s, alphas, betas, _ = synth_inputs(p, nkeys, seed=n)
    o = orc.gen(p, s0s, alphas, None if scheme == "grotto" else betas)
    oc, ooc = o if scheme == "halftree" else (o, None)
    cws, ocws = T(oc, dev), (None if ooc is None else T(ooc, dev))
    for party in (0, 1):
        got = ctx.eval_all(party, T(s0s[:, party], dev), cws, ocws)
        want = orc.evalall(p, party, s0s[:, party], oc, ooc, threads=8)
        assert np.array_equal(N(got, want.dtype), want), party
    # leaf sub-ranges in whole work units (multi-GPU subtree sharding)
    g = ctx.granule()
    assert g == 1 << min(n, 16 if scheme == "dcf" else 17)
    if (1 << n) > g and scheme != "grotto":
        full = orc.evalall(p, 0, s0s[:, 0], oc, ooc, threads=8)
        for b, cnt in ((g, g), (0, g), (g, 0)):
            got = ctx.eval_all(0, T(s0s[:, 0], dev), cws, ocws, leaf_begin=b, leaf_count=cnt)
            assert np.array_equal(N(got), full[:, b:b + (cnt or (1 << n) - b)])
    if scheme == "grotto":
        t = ctx.grotto_expand(1, T(s0s[:, 1], dev), cws)
        assert np.array_equal(N(t, np.uint8), orc.grotto_expand(p, 1, s0s[:, 1], oc, threads=8))
        if n <= 20:
            pt = ctx.grotto_preprocess(1, T(s0s[:, 1], dev), cws)
            want = orc.grotto_preprocess(p, 1, s0s[:, 1], oc)
            assert np.array_equal(N(pt, np.uint8), want)
            xs = [0, (1 << n) - 1, 7, (1 << n) - 2, 100][:nkeys]
            assert np.array_equal(N(ctx.grotto_lookup(pt[:len(xs)], xs), np.uint8),
                                  orc.grotto_lookup(p, want[:len(xs)], xs))


# ---- host-buffer entry points (what a CPU caller of the reference binds) ---------------------------------------------

def test_host_entry_points(dev, orc):
    p = Params(scheme="dpf", in_bits=32)
    ctx = mkctx(p)
    ctx.reserve_host(1000)  # force several chunks
    s0s, alphas, betas, xs = synth_inputs(p, 4500, seed=8)
    cws = ctx.gen(torch.from_numpy(s0s.view(np.int32)), alphas, torch.from_numpy(betas.view(np.int32)))
    assert cws.device.type == "cpu"
    oc = orc.gen(p, s0s, alphas, betas, threads=8)
    assert np.array_equal(N(cws), oc)
    ys = ctx.eval(1, torch.from_numpy(np.ascontiguousarray(s0s[:, 1]).view(np.int32)), cws, xs)
    assert ys.device.type == "cpu"
    assert np.array_equal(N(ys), orc.eval(p, 1, s0s[:, 1], oc, xs, threads=8))
    p2 = Params(scheme="dpf", in_bits=19, group="u64")
    ctx2 = mkctx(p2)
    s0s, alphas, betas, xs = synth_inputs(p2, 3, seed=9)
    oc = orc.gen(p2, s0s, alphas, betas)
    ya = ctx2.eval_all(0, torch.from_numpy(np.ascontiguousarray(s0s[:, 0]).view(np.int32)),
                       torch.from_numpy(oc.view(np.int32)))
    assert ya.device.type == "cpu" and ya.shape == (3, 1 << 19, 4)
    assert np.array_equal(N(ya), orc.evalall(p2, 0, s0s[:, 0], oc, threads=8))


@pytest.mark.parametrize("scheme,n,group,nkeys,cap,set_mb", [
    ("dpf", 10, "u64", 7, 0, None),       # every key in one launch
    ("dpf", 10, "bytes", 7, 3, None),     # groups of 3, 3, 1 keys
    ("halftree", 9, "u32", 5, 2, None),   # ... with per-key ocws
    ("grotto", 11, "bytes", 5, 2, None),  # ... byte leaves + the scan
    ("dcf", 12, "u128", 4, 0, None),
    ("dpf", 18, "u64", 3, 0, "1"),        # 1 MiB sets: one key at a time, 4 leaf ranges each
    ("dpf", 17, "bytes", 3, 0, "4"),      # 4 MiB sets: two whole keys per launch, then one
])
def test_eval_all_host_geometries(dev, orc, monkeypatch, scheme, n, group, nkeys, cap, set_mb):
    """fssb200_eval_all_host: whole keys per launch for small domains, leaf ranges of one key for large ones."""
    p = Params(scheme=scheme, in_bits=n, group=group, hash_key=HASH_KEY_BENCH)
    ctx = mkctx(p)
    ctx.reserve_host(cap)
    if set_mb:
        monkeypatch.setenv("FSSB200_ALL_SET_MB", set_mb)
    s0s, alphas, betas, _ = synth_inputs(p, nkeys, seed=40 + n)
    o = orc.gen(p, s0s, alphas, None if scheme == "grotto" else betas)
    oc, ooc = o if scheme == "halftree" else (o, None)
    H = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32))
    for party in (0, 1):
        got = ctx.eval_all(party, H(s0s[:, party]), H(oc), None if ooc is None else H(ooc))
        assert got.device.type == "cpu"
        want = orc.evalall(p, party, s0s[:, party], oc, ooc, threads=8)
        assert np.array_equal(N(got, want.dtype), want), party
    g = ctx.granule()
    if (1 << n) > g and scheme != "grotto":
        got = ctx.eval_all(0, H(s0s[:, 0]), H(oc), None if ooc is None else H(ooc), leaf_begin=g, leaf_count=g)
        assert np.array_equal(N(got, want.dtype), orc.evalall(p, 0, s0s[:, 0], oc, ooc, threads=8)[:, g:2 * g])


# ---- packed rows (compact key format) and the packing host path -----------------------------------------------------

@pytest.mark.parametrize("scheme,n,group,prg,nkeys", [
    ("dpf", 32, "bytes", "aes128_mmo", 20000), ("halftree", 20, "u64", "aes128_mmo", 9001),
    ("dpf", 64, "u128", "chacha", 8200), ("dpf", 128, "bytes", "aes128_mmo", 300), ("dpf", 5, "u32", "aes128_mmo", 33),
    ("halftree", 32, "bytes", "chacha", 1), ("dpf", 3, "bytes", "aes128_mmo", 8193), ("halftree", 1, "u32", "aes128_mmo", 70),
])
def test_packed_rows(dev, orc, scheme, n, group, prg, nkeys):
    p = Params(scheme=scheme, in_bits=n, group=group, prg=prg, hash_key=HASH_KEY_BENCH)
    s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=7 * n + nkeys)
    xs[0] = (1 << n) - 1
    o = orc.gen(p, s0s, alphas, betas, threads=8)
    oc, ooc = o if scheme == "halftree" else (o, None)
    want = [orc.eval(p, party, s0s[:, party], oc, xs, ooc, threads=8) for party in (0, 1)]
    ctx = mkctx(p)
    cws_h = torch.from_numpy(oc.view(np.int32))
    ocws_d = None if ooc is None else T(ooc, dev)
    # the format itself: ncw 16-byte s entries + 16 bytes of flag bits (bit i = byte 16 of entry i != 0)
    rows = ctx.pack_rows(cws_h)
    ncw = p.ncw
    assert rows.shape == (nkeys, ncw * 16 + 16) and ctx.packed_row_bytes() == ncw * 16 + 16
    r = rows.numpy()
    raw = oc.view(np.uint8).reshape(nkeys, ncw, 32)
    assert np.array_equal(r[:, :ncw * 16].reshape(nkeys, ncw, 16), raw[:, :, :16])
    bits = (raw[:, :min(ncw, 128), 16] != 0)
    flags = np.zeros((nkeys, 128), dtype=bool)
    flags[:, :bits.shape[1]] = bits
    assert np.array_equal(r[:, ncw * 16:], np.packbits(flags, axis=1, bitorder="little"))
    # device evaluation on packed rows
    rows_d = rows.to(dev)
    for party in (0, 1):
        ys = ctx.eval_packed(party, T(s0s[:, party], dev), rows_d, xs, ocws_d)
        assert np.array_equal(N(ys), want[party]), party
    # host entry point: adaptive pipeline (0), reference layout only (1), staged chunks only (2); several chunks
    seeds_h = torch.from_numpy(np.ascontiguousarray(s0s[:, 1]).view(np.int32))
    ocws_h = None if ooc is None else torch.from_numpy(ooc.view(np.int32))
    ctx_h = mkctx(p)
    ctx_h.reserve_host(2048)
    for mode in (0, 1, 2):
        ctx_h.set_host_mode(mode)
        for pin in (False, True):
            a = [seeds_h, cws_h, ocws_h]
            if pin:
                a = [None if v is None else v.pin_memory() for v in a]
            ys = ctx_h.eval(1, a[0], a[1], xs, a[2])
            assert ys.device.type == "cpu" and np.array_equal(N(ys), want[1]), (mode, pin)


def test_packed_rows_rejects_dcf(dev):
    ctx = mkctx(Params(scheme="dcf", in_bits=16))
    assert ctx.packed_row_bytes() == 0
    with pytest.raises(ValueError):
        ctx.pack_rows(torch.zeros((4, 17, 8), dtype=torch.int32))


# ---- full-size configurations through size-independent properties ---------------------------------------------------------

def test_c2_full_size_reconstruction(dev, orc):
    """BASELINE config 2: 2^22 DPF keys, n = 32, Bytes, AES-128 MMO: y0 ^ y1 == (x == alpha ? beta : 0)."""
    p = Params(scheme="dpf", in_bits=32)
    ctx = mkctx(p)
    k = 1 << 22
    g = torch.Generator(device=dev).manual_seed(42)
    s0s = torch.randint(-2 ** 31, 2 ** 31, (k, 2, 4), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    betas = torch.randint(-2 ** 31, 2 ** 31, (k, 4), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    s0s[:, :, 3] &= ~1
    betas[:, 3] &= ~1
    alphas = torch.randint(-2 ** 31, 2 ** 31, (k,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    xs = torch.randint(-2 ** 31, 2 ** 31, (k,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    xs[::16] = alphas[::16]
    cws = ctx.gen(s0s, alphas, betas)
    y0 = ctx.eval(0, s0s[:, 0].contiguous(), cws, xs)
    y1 = ctx.eval(1, s0s[:, 1].contiguous(), cws, xs)
    hit = (xs == alphas).unsqueeze(1)
    assert int(hit.sum()) >= k // 16
    assert torch.equal(y0 ^ y1, torch.where(hit, betas, torch.zeros_like(betas)))
    # a slice against the oracle (keys generated on the GPU, evaluated on the CPU)
    sl = slice(123456, 123456 + 2048)
    want = orc.eval(p, 0, N(s0s[sl, 0]), N(cws[sl]), N(xs[sl]).astype(np.uint32), threads=8)
    assert np.array_equal(N(y0[sl]), want)


def _u127_add(a, b):
    """(a + b) mod 2^127 on (N,4) int32 tensors in the Uint<u128,2^127> wire format (uint.cuh:58-62,76-81)."""
    m = 0xFFFFFFFF
    a64 = [a[:, i].to(torch.int64) & m for i in range(4)]
    b64 = [b[:, i].to(torch.int64) & m for i in range(4)]
    a64[3], b64[3] = a64[3] >> 1, b64[3] >> 1
    out, carry = [], torch.zeros_like(a64[0])
    for i in range(4):
        s = a64[i] + b64[i] + carry
        out.append(s & m)
        carry = s >> 32
    out[3] = (out[3] & 0x7FFFFFFF) << 1
    return torch.stack([torch.where(o >= 2 ** 31, o - 2 ** 32, o).to(torch.int32) for o in out], dim=1)


def test_c3_full_size_reconstruction(dev, orc):
    """BASELINE config 3 at full size: DCF, n = 64, Uint<u128, 2^127>, Aes128Mmo<4>, 2^22 keys (8.7 GB of keys):
    y0 + y1 == (x < alpha ? beta : 0)."""
    p = Params(scheme="dcf", in_bits=64, group="u128")
    ctx = mkctx(p)
    k = 1 << 22
    g = torch.Generator(device=dev).manual_seed(43)
    s0s = torch.randint(-2 ** 31, 2 ** 31, (k, 2, 4), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    betas = torch.randint(-2 ** 31, 2 ** 31, (k, 4), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    s0s[:, :, 3] &= ~1
    betas[:, 3] &= ~1
    # bit patterns with the top bit set compare as unsigned after flipping the sign bit
    alphas = torch.randint(-2 ** 63, 2 ** 63 - 1, (k,), dtype=torch.int64, device=dev, generator=g)
    xs = torch.randint(-2 ** 63, 2 ** 63 - 1, (k,), dtype=torch.int64, device=dev, generator=g)
    xs[::16] = alphas[::16]
    xs[1::16] = alphas[1::16] - 1
    cws = ctx.gen(s0s, alphas, betas)
    y0 = ctx.eval(0, s0s[:, 0].contiguous(), cws, xs)
    y1 = ctx.eval(1, s0s[:, 1].contiguous(), cws, xs)
    sign = torch.tensor(-2 ** 63, dtype=torch.int64, device=dev)
    lt = ((xs ^ sign) < (alphas ^ sign)).unsqueeze(1)   # unsigned x < alpha
    betas_c = betas.clone()
    tot = _u127_add(y0, y1)
    assert torch.equal(tot, torch.where(lt, betas_c, torch.zeros_like(betas_c)))
    sl = slice(777, 777 + 1024)
    wanto = orc.eval(p, 1, N(s0s[sl, 1]), N(cws[sl]), N(xs[sl], np.uint64), threads=8)
    assert np.array_equal(N(y1[sl]), wanto)


def test_c4_full_domain_n28(dev, golden):
    """BASELINE config 4 domain: DPF EvalAll n = 28 (2^28 leaves, 4 GiB) of one key, against the SHA-256 of the
    reference's own EvalAll output (tests/golden), plus reconstruction over the whole domain."""
    c = golden.by_name("c4_dpf_n28_bytes_aes")
    p = c.p
    ctx = mkctx(p)
    s0s, cws = T(c["s0s"], dev), T(c["cws"], dev)
    y0 = ctx.eval_all(0, s0s[:1, 0].contiguous(), cws[:1])
    idx = torch.tensor(c.meta["all_sample_idx"], device=dev)
    assert np.array_equal(N(y0[0, idx]), c["all0_sample"][0])
    y1 = ctx.eval_all(1, s0s[:1, 1].contiguous(), cws[:1])
    assert np.array_equal(N(y1[0, idx]), c["all1_sample"][0])
    tot = y0[0] ^ y1[0]
    nz = tot.ne(0).any(dim=1).nonzero().flatten()
    assert nz.tolist() == [c.alphas[0]]
    assert np.array_equal(N(tot[c.alphas[0]]), c["betas"][0])
    del y1, tot
    h = hashlib.sha256()
    host = y0[0].cpu().numpy()
    h.update(host.tobytes())
    assert h.hexdigest() == c.meta["all0_sha256"][0]


# ---- the reference binding's own integration tests, restated (test/test_dpf_integration.py, test_dcf_integration.py) --

@pytest.mark.parametrize("cls,kw", [("Dpf", {}), ("Dcf", {"pred": "lt"}), ("Dcf", {"pred": "gt"})])
def test_fss_crypto_dropin(dev, cls, kw):
    import fss_crypto
    sch = getattr(fss_crypto, cls)(in_bits=16, group="bytes", prg="chacha", **kw)
    s0s = torch.randint(-2 ** 31, 2 ** 31, (2, 4), dtype=torch.int64).to(torch.int32)
    beta = torch.tensor([0, 0, 0, 604], dtype=torch.int32)
    cws = sch.gen(s0s, alpha=107, beta=beta)
    assert cws.shape == (17, 8) and cws.dtype == torch.int32 and cws.device.type == "cpu"
    out = sch.eval(party=0, s0=s0s[0], cws=cws, x=50)
    assert out.shape == (4,) and out.dtype == torch.int32 and out.device.type == "cpu"
    with pytest.raises(ValueError, match="x must be"):
        sch.eval(party=0, s0=s0s[0], cws=cws, x=2 ** 16)
    outc = sch.eval(party=0, s0=s0s[0].to(dev), cws=cws.to(dev), x=50)
    assert outc.shape == (4,) and outc.device.type == "cuda" and torch.equal(outc.cpu(), out)
    if cls == "Dpf":
        ya = sch.eval_all(party=0, s0=s0s[0], cws=cws)
        assert ya.shape == (2 ** 16, 4) and ya.dtype == torch.int32 and ya.device.type == "cpu"
        yb = sch.eval_all(party=1, s0=s0s[1], cws=cws)
        tot = ya ^ yb
        assert torch.equal(tot[107], beta) and int(tot.ne(0).any(dim=1).sum()) == 1
    # reconstruction through the single-key API (src/dpf_test.cu:58-78, dcf_test.cu:91-135)
    for x in (0, 106, 107, 108, 2 ** 16 - 1):
        y = sch.eval(0, s0s[0], cws, x) ^ sch.eval(1, s0s[1], cws, x)
        if cls == "Dpf":
            want = beta if x == 107 else torch.zeros_like(beta)
        else:
            hit = x < 107 if kw["pred"] == "lt" else x > 107
            want = beta if hit else torch.zeros_like(beta)
        assert torch.equal(y, want), (cls, kw, x)


def test_error_codes_on_device(dev):
    import fss_b200
    from fss_b200 import _lib as L
    ctx = fss_b200.Context("dpf", 32)
    h = ctx.handle(0)
    buf = torch.zeros(4096, dtype=torch.int32, device=dev)
    p = C.c_void_p(buf.data_ptr())
    assert L.lib.fssb200_dcf_eval(h, 0, p, p, p, p, 1, None) == L.E_SCHEME
    assert L.lib.fssb200_dpf_eval(h, 2, p, p, p, p, 1, None) == L.E_INVAL
    assert L.lib.fssb200_dpf_eval(h, 0, C.c_void_p(buf.data_ptr() + 4), p, p, p, 1, None) == L.E_ALIGN
    assert L.lib.fssb200_dpf_eval(h, 0, None, p, p, p, 1, None) == L.E_INVAL
    assert L.lib.fssb200_dpf_eval(h, 0, p, p, p, p, 0, None) == 0       # empty batch is a no-op
    small = fss_b200.Context("dpf", 20)
    hs = small.handle(0)
    assert L.lib.fssb200_eval_all(hs, 0, p, p, None, p, 1, 5, 0, None) == L.E_RANGE       # not unit aligned
    assert L.lib.fssb200_eval_all(hs, 0, p, p, None, p, 1, 1 << 20, 0, None) == L.E_RANGE  # outside the domain
    gr = 1 << 17                                                                            # begin + count wraps uint64
    assert L.lib.fssb200_eval_all(hs, 0, p, p, None, p, 1, gr, 2 ** 64 - gr, None) == L.E_RANGE
    assert L.lib.fssb200_eval_all_host(hs, 0, p, p, None, p, 1, gr, 2 ** 64 - gr) == L.E_RANGE
    assert L.lib.fssb200_eval_host(h, 2, p, p, None, p, p, 1) == L.E_INVAL
    assert L.lib.fssb200_ctx_set_host_mode(h, 4) == L.E_INVAL
    g = fss_b200.Context("grotto", 10)
    assert L.lib.fssb200_eval(g.handle(0), 0, p, p, None, p, p, 1, None) == L.E_SCHEME


def test_smoke_entry(dev):
    import __graft_entry__ as ge
    ge.smoke()


def test_cpp_header_shim(dev, tmp_path):
    """The header-only C++ surface (include/fss/*.cuh -> C ABI), compiled with g++ -std=c++20 and run: the
    reference samples' flows, the survey KATs, the fss::gpu::* entry points (tests/cpp/shim_sample.cpp)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "shim_sample")
    subprocess.run(["g++", "-std=c++20", "-O1", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(root, "tests", "cpp", "shim_sample.cpp"), "-o", exe, "-L", os.path.join(root, "fss_b200"),
                    "-lfssb200", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(root, "fss_b200")],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr
