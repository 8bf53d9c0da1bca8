"""Seeded random sweep over the parameter space of the C ABI against the oracle (`pytest -m gpu`).

The parametrised parity tests pin chosen corners; this file draws the corners: scheme x in_bits (1..128) x In width x
group (every Uint width, wrap-around / power-of-two / odd / near-top moduli, Bytes) x PRG x predicate x ragged batch size x
party x entry point (VDPF included; device arrays, host arrays in every host mode with pinned or pageable memory and a random chunk size,
packed rows, the level-major layout, full-domain evaluation with a leaf sub-range).  Bar: bit-exact.

FSSB200_FUZZ_SEED / FSSB200_FUZZ_CASES widen it (tools/gpu_session.sh fuzz runs a few thousand cases)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import HASH_KEY_BENCH, Params
from test_gpu_parity import N, T, mkctx

pytestmark = pytest.mark.gpu

SEED = int(os.environ.get("FSSB200_FUZZ_SEED", "20260"))
CASES = int(os.environ.get("FSSB200_FUZZ_CASES", "96"))
WIDTH = {"u8": 8, "u16": 16, "u32": 32, "u64": 64, "u128": 127}


def draw_params(r: random.Random) -> Params:
    scheme = r.choice(["dpf", "dpf", "dcf", "dcf", "halftree", "grotto", "vdpf"])
    n = r.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 127, 128, r.randint(1, 128), r.randint(1, 40)])
    widths = [b for b in (1, 2, 4, 8, 16) if 8 * b >= n]
    in_bytes = r.choice(widths[:2] + [widths[0]])
    group, mod = "bytes", 0
    if scheme != "grotto" and r.random() < 0.75:
        group = r.choice(list(WIDTH))
        w = WIDTH[group]
        kind = r.randrange(5)
        if kind == 0:
            mod = 0                                    # wrap-around (u128: 2^127, uint.cuh:58-62)
        elif kind == 1:
            mod = 1 << r.randint(1, w - 1)             # power of two below the width
        elif kind == 2:
            mod = r.randrange(3, 1 << w) | 1           # any odd modulus
        elif kind == 3:
            mod = (1 << w) - r.choice([1, 3, 5, 59, 189])  # just below the top
        else:
            mod = r.randrange(2, 1 << min(w, 20))      # small, even or odd
    hashes = (r.choice(["blake3", "sha256"]), r.choice(["blake3", "sha256"])) if scheme == "vdpf" else ("blake3", "blake3")
    return Params(scheme=scheme, in_bits=n, group=group, mod=mod, prg=r.choice(["aes128_mmo", "chacha"]),
                  pred=r.choice(["lt", "gt"]) if scheme == "dcf" else "lt", hash_key=HASH_KEY_BENCH, in_bytes=in_bytes,
                  hash=hashes)


def draw_inputs(r: random.Random, p: Params, nkeys: int):
    rng = np.random.default_rng(r.getrandbits(32))
    s0s = rng.integers(0, 2 ** 32, size=(nkeys, 2, 4), dtype=np.uint64).astype(np.uint32)
    s0s[:, :, 3] &= 0xFFFFFFFE
    betas = rng.integers(0, 2 ** 32, size=(nkeys, 4), dtype=np.uint64).astype(np.uint32)
    betas[:, 3] &= 0xFFFFFFFE
    top = (1 << p.in_bits) - 1
    alphas = [r.getrandbits(p.in_bits) for _ in range(nkeys)]
    xs = [r.getrandbits(p.in_bits) for _ in range(nkeys)]
    for i in range(nkeys):                             # the branches: x == alpha, neighbours, the domain's ends
        k = r.randrange(12)
        if k == 0:
            xs[i] = alphas[i]
        elif k == 1:
            xs[i] = (alphas[i] + 1) & top
        elif k == 2:
            xs[i] = (alphas[i] - 1) & top
        elif k == 3:
            xs[i] = r.choice([0, top])
        elif k == 4:
            alphas[i] = r.choice([0, top])
    return s0s, alphas, betas, xs


def host(a):
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a)


def run_vdpf(r, orc, dev, tag, p, s0s, alphas, betas, xs):
    """vdpf.cuh Gen / Eval (+ the proof tuples) / BatchProve inputs, device and host arrays."""
    import fss_b200
    ctx = fss_b200.Context("vdpf", p.in_bits, p.group, mod=p.mod, prg=p.prg, prg_key=p.prg_key, in_bytes=p.in_bytes,
                           hash_iv=bytes(p.hash_iv), hash=p.hash)
    want = orc.vdpf_gen(p, s0s, alphas, betas, threads=8)
    got = ctx.vdpf_gen(T(s0s, dev), alphas, T(betas, dev))
    for u, v in zip(want, got):
        assert np.array_equal(u, N(v).view(u.dtype).reshape(u.shape)), (tag, "vdpf gen")
    cws, cs, ocws, _ = got
    party = r.randrange(2)
    seeds = np.ascontiguousarray(s0s[:, party])
    wy, wp = orc.vdpf_eval(p, party, seeds, want[0], want[1], want[2], xs, threads=8)
    ys, pt = ctx.vdpf_eval(party, T(seeds, dev), cws, cs, ocws, xs)
    assert np.array_equal(N(ys), wy) and np.array_equal(N(pt), wp), (tag, "vdpf eval")
    ctx.reserve_host(r.choice([0, 1, 50, 1000]))
    hy, hp = ctx.vdpf_eval(party, host(seeds), cws.cpu(), cs.cpu(), ocws.cpu(), xs)
    assert hy.device.type == "cpu" and np.array_equal(N(hy), wy) and np.array_equal(N(hp), wp), (tag, "vdpf eval host")
    if p.in_bits <= 12:
        k = min(len(xs), 3)
        ya, pa = ctx.vdpf_eval_all(party, T(seeds[:k], dev), cws[:k], cs[:k], ocws[:k])
        way, wpa = orc.vdpf_evalall(p, party, seeds[:k], want[0][:k], want[1][:k], want[2][:k], threads=4)
        assert np.array_equal(N(ya), way) and np.array_equal(N(pa), wpa), (tag, "vdpf evalall")


def run_case(r: random.Random, orc, dev, tag):
    p = draw_params(r)
    nkeys = r.choice([1, 2, 31, 32, 33, 511, 513, r.randint(1, 300), r.randint(1, 3000)])
    if r.randrange(16) == 0:
        nkeys = r.randint(8192, 20000)                 # the pipelined host path (fssb200_eval_host from 8192 keys on)
    if p.scheme == "grotto":
        nkeys = min(nkeys, 700)
    s0s, alphas, betas, xs = draw_inputs(r, p, nkeys)
    if p.scheme == "vdpf":
        run_vdpf(r, orc, dev, tag, p, s0s, alphas, betas, xs)
        return p
    ctx = mkctx(p)
    gb = None if p.scheme == "grotto" else betas
    o = orc.gen(p, s0s, alphas, gb, threads=8)
    oc, ooc = o if p.scheme == "halftree" else (o, None)

    got = ctx.gen(T(s0s, dev), alphas, None if gb is None else T(gb, dev))
    gc, goc = got if p.scheme == "halftree" else (got, None)
    assert np.array_equal(N(gc), oc), (tag, "gen")
    if goc is not None:
        assert np.array_equal(N(goc), ooc), (tag, "gen ocw")

    party = r.randrange(2)
    seeds = np.ascontiguousarray(s0s[:, party])
    ocw_d = None if ooc is None else T(ooc, dev)

    if p.scheme == "grotto":
        w = [N(ctx.grotto_walk(b, T(s0s[:, b], dev), T(oc, dev), xs), np.uint8) for b in (0, 1)]
        want = np.array([1 if a <= x else 0 for a, x in zip(alphas, xs)], np.uint8)
        assert np.array_equal(w[0] ^ w[1], want), (tag, "walk")
        if p.in_bits <= 12:
            k = min(nkeys, 16)
            ya = N(ctx.eval_all(party, T(seeds[:k], dev), T(oc[:k], dev)), np.uint8)
            assert np.array_equal(ya, orc.evalall(p, party, seeds[:k], oc[:k], threads=8)), (tag, "grotto evalall")
        return p

    want = orc.eval(p, party, seeds, oc, xs, ooc, threads=8)
    assert np.array_equal(N(ctx.eval(party, T(seeds, dev), T(oc, dev), xs, ocw_d)), want), (tag, "eval device")

    # host arrays: a random host mode, pinned or pageable, a random chunk size
    mode, pin = r.randrange(3), r.random() < 0.5
    ctx.set_host_mode(mode)
    ctx.reserve_host(r.choice([0, 0, 1, 7, 100, 1000, 4096]))
    a = [host(seeds), host(oc), None if ooc is None else host(ooc)]
    if pin:
        a = [None if v is None else v.pin_memory() for v in a]
    ys = ctx.eval(party, a[0], a[1], xs, a[2])
    assert ys.device.type == "cpu" and np.array_equal(N(ys), want), (tag, "eval host", mode, pin)
    ctx.set_host_mode(0)

    if ctx.packed_row_bytes():
        rows = ctx.pack_rows(host(oc)).to(dev)
        assert np.array_equal(N(ctx.eval_packed(party, T(seeds, dev), rows, xs, ocw_d)), want), (tag, "eval packed")

    lay = ctx.relayout(T(oc, dev))
    assert np.array_equal(N(ctx.eval_levelmajor(party, T(seeds, dev), lay, xs, ocw_d)), want), (tag, "eval level-major")

    if p.in_bits <= 13 or (p.in_bits <= 18 and nkeys <= 3):
        k = min(nkeys, 4)
        g = ctx.granule()
        units = (1 << p.in_bits) // g
        b = r.randrange(units) * g
        cnt = r.randint(1, units - b // g) * g
        ya = ctx.eval_all(party, T(seeds[:k], dev), T(oc[:k], dev), None if ooc is None else T(ooc[:k], dev),
                          leaf_begin=b, leaf_count=cnt)
        full = orc.evalall(p, party, seeds[:k], oc[:k], None if ooc is None else ooc[:k], threads=8)
        assert np.array_equal(N(ya), full[:, b:b + cnt]), (tag, "evalall", b, cnt)
    return p


@pytest.mark.parametrize("block", range(4))
def test_random_parameter_sweep(orc, block):
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    dev = torch.device("cuda:0")
    per = (CASES + 3) // 4
    seen = set()
    for i in range(per):
        r = random.Random(SEED * 1000003 + block * 100003 + i)
        p = run_case(r, orc, dev, (SEED, block, i))
        seen.add((p.scheme, p.prg))
    if per >= 16:
        assert len(seen) >= 6
