"""CPU check of the CUDA kernels' per-thread bodies.

tests/host_emul/host_emul.cpp compiles the `__host__ __device__` code of fss_b200/csrc/{aes,prg,group,
schemes}.cuh for the host (same T-table image, PRMT address formation, lane replication, packed-node
correction, group arithmetic) and this test compares it bit for bit with the oracle.  It is how the
kernel logic is validated in the GPU-less build container; the `-m gpu` tests then cover the real
kernels through the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import HASH_KEY_BENCH, Params, _vp, pack_ints, synth_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "host_emul.cpp")
LIB = os.path.join(HERE, "host_emul", "_host_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "fss_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("common.cuh", "aes.cuh", "blake3.cuh", "prg.cuh", "group.cuh",
                                                          "schemes.cuh", "sha256.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++", "-w", "-I/usr/local/cuda/include",
                        SRC, "-o", LIB], check=True)
    return C.CDLL(LIB)


def emul_gen(emu, p, s0s, alphas, betas):
    k = len(s0s)
    cws, ocws = np.zeros((k, p.ncw, 8), np.uint32), np.zeros((k, 4), np.uint32)
    cp, al = p.c(), pack_ints(alphas, p.in_bytes)
    emu.emul_gen(C.byref(cp), C.c_size_t(k), _vp(np.ascontiguousarray(s0s)), _vp(al), _vp(betas), _vp(cws), _vp(ocws))
    return cws, ocws


def emul_eval(emu, p, party, seeds, cws, xs, ocws=None, lm=None):
    k = len(seeds)
    ys, cp, xb = np.zeros((k, 4), np.uint32), p.c(), pack_ints(xs, p.in_bytes)
    seeds = np.ascontiguousarray(seeds)
    if lm is None:
        emu.emul_eval(C.byref(cp), party, C.c_size_t(k), _vp(seeds), _vp(cws), _vp(ocws), _vp(xb), _vp(ys), 0, None,
                      None, None, None)
    else:
        emu.emul_eval(C.byref(cp), party, C.c_size_t(k), _vp(seeds), None, _vp(ocws), _vp(xb), _vp(ys), 1,
                      _vp(lm[0]), _vp(lm[1]), _vp(lm[2]), _vp(lm[3]))
    return ys


def emul_all(emu, p, party, seeds, cws, ocws=None, lb=0, cnt=0):
    k, n = len(seeds), cnt or ((1 << p.in_bits) - lb)
    ys = np.zeros((k, n), np.uint8) if p.scheme == "grotto" else np.zeros((k, n, 4), np.uint32)
    cp = p.c()
    rc = emu.emul_evalall(C.byref(cp), party, C.c_size_t(k), _vp(np.ascontiguousarray(seeds)), _vp(cws), _vp(ocws),
                          _vp(ys), C.c_uint64(lb), C.c_uint64(cnt))
    assert rc == 0
    return ys


@pytest.mark.parametrize("prg", ["aes128_mmo", "chacha"])
def test_prg_blocks(emu, orc, golden, prg):
    seeds = golden.arrays["prg/seeds"]
    for mul in (1, 2, 4):
        p = Params(prg=prg)
        out = np.zeros((len(seeds), mul, 4), np.uint32)
        cp = p.c()
        emu.emul_prg_gen(C.byref(cp), mul, C.c_size_t(len(seeds)), _vp(seeds), _vp(out))
        assert np.array_equal(out, golden.arrays[f"prg/{prg}_{mul}"])   # reference-generated
        assert np.array_equal(out, orc.prg_gen(p, mul, seeds))


MODS = {"bytes": [0], "u8": [0, 251], "u16": [65521], "u32": [0, 4294967291, 7], "u64": [0, 18446744073709551557],
        "u128": [1 << 127, (1 << 127) - 1, 5, (1 << 64) + 13]}


@pytest.mark.parametrize("prg", ["aes128_mmo", "chacha"])
@pytest.mark.parametrize("scheme", ["dpf", "dcf", "halftree", "grotto"])
def test_kernel_bodies_match_oracle(emu, orc, prg, scheme):
    for n in (1, 2, 5, 8, 12, 32, 33, 64, 65, 128):
        for group, mods in MODS.items():
            if scheme == "grotto" and group != "bytes":
                continue
            if n not in (8, 32, 64) and group not in ("bytes", "u64", "u128"):
                continue
            for mod in mods:
                for pred in (("lt", "gt") if scheme == "dcf" and n in (8, 64) else ("lt",)):
                    p = Params(scheme=scheme, in_bits=n, group=group, mod=mod, prg=prg, pred=pred,
                               hash_key=HASH_KEY_BENCH)
                    k = 40  # > 32 so that every table replica (lane) is exercised
                    s0s, alphas, betas, xs = synth_inputs(p, k, seed=n + len(group))
                    xs[1], xs[2], alphas[3], xs[3], alphas[4], xs[4] = 0, (1 << n) - 1, 0, 0, (1 << n) - 1, (1 << n) - 1
                    o = orc.gen(p, s0s, alphas, betas)
                    oc, ooc = o if scheme == "halftree" else (o, None)
                    ec, eoc = emul_gen(emu, p, s0s, alphas, None if scheme == "grotto" else betas)
                    tag = (scheme, n, group, hex(mod), prg, pred)
                    assert np.array_equal(oc, ec), ("gen",) + tag
                    assert ooc is None or np.array_equal(ooc, eoc), ("gen ocw",) + tag
                    if scheme != "grotto":
                        lm = orc.relayout(p, oc)
                        for party in (0, 1):
                            want = orc.eval(p, party, s0s[:, party], oc, xs, ooc)
                            assert np.array_equal(want, emul_eval(emu, p, party, s0s[:, party], oc, xs, ooc)), tag
                            assert np.array_equal(want, emul_eval(emu, p, party, s0s[:, party], oc, xs, ooc, lm)), tag
                    if n <= 12 and scheme != "dcf":
                        for party in (0, 1):
                            so = None if ooc is None else ooc[:3]
                            want = (orc.grotto_expand(p, party, s0s[:3, party], oc[:3]) if scheme == "grotto" else
                                    orc.evalall(p, party, s0s[:3, party], oc[:3], so))
                            assert np.array_equal(want, emul_all(emu, p, party, s0s[:3, party], oc[:3], so)), tag


def test_golden_through_kernel_bodies(emu, golden):
    """Reference-generated fixtures replayed through the kernel bodies (no oracle in the loop)."""
    for c in golden.cases:
        p = c.p
        if p.scheme == "grotto":
            continue
        ocws = c["ocws"] if p.scheme == "halftree" else None
        for party in (0, 1):
            got = emul_eval(emu, p, party, c["s0s"][:, party], c["cws"], c.xs, ocws)
            assert np.array_equal(got, c[f"ys{party}"]), (c.name, party)
        ec, eoc = emul_gen(emu, p, c["s0s"], c.alphas, c.betas)
        assert np.array_equal(c.masked_cws(ec), c.masked_cws(c["cws"])), c.name


def emul_walk(emu, p, party, seeds, cws, xs):
    k = len(seeds)
    ys, cp, xb = np.zeros(k, np.uint8), p.c(), pack_ints(xs, p.in_bytes)
    rc = emu.emul_grotto_walk(C.byref(cp), party, C.c_size_t(k), _vp(np.ascontiguousarray(seeds)), _vp(cws), _vp(xb),
                              _vp(ys))
    assert rc == 0
    return ys


@pytest.mark.parametrize("prg", ["aes128_mmo", "chacha"])
def test_grotto_walk_reconstructs_like_the_reference(emu, orc, prg):
    """The O(n) Grotto walk returns a different SHARE than GrottoDcf::Eval but the same SECRET:
    share0 ^ share1 == 1[alpha <= x] == reference Preprocess + Eval reconstructed (grotto_dcf.cuh:94-135)."""
    for n, in_bytes in ((1, 1), (2, 4), (5, 1), (8, 1), (8, 4), (10, 2), (16, 2), (32, 4), (33, 8), (64, 8), (100, 16),
                        (128, 16)):
        p = Params(scheme="grotto", in_bits=n, prg=prg, in_bytes=in_bytes)
        k = 48
        s0s, alphas, _, xs = synth_inputs(p, k, seed=100 + n)
        top = (1 << n) - 1
        xs[0], xs[1], alphas[2], xs[2], alphas[3], xs[3] = 0, top, 0, 0, top, top
        xs[4], alphas[5], xs[5] = alphas[4], top, max(0, top - 1)
        cws = orc.gen(p, s0s, alphas, None)
        w0 = emul_walk(emu, p, 0, s0s[:, 0], cws, xs)
        w1 = emul_walk(emu, p, 1, s0s[:, 1], cws, xs)
        want = np.array([1 if int(a) <= int(x) else 0 for a, x in zip(alphas, xs)], np.uint8)
        assert np.array_equal(w0 ^ w1, want), (n, in_bytes, prg)
        if n <= 10:  # every point of the domain against the reference's parity-tree shares, reconstructed
            m = min(k, 6)
            for kk in range(m):
                allx = list(range(1 << n))
                rep = lambda a: np.repeat(a[kk:kk + 1], len(allx), axis=0)
                a0 = emul_walk(emu, p, 0, rep(s0s[:, 0]), rep(cws), allx)
                a1 = emul_walk(emu, p, 1, rep(s0s[:, 1]), rep(cws), allx)
                r0 = orc.evalall(p, 0, s0s[kk:kk + 1, 0], cws[kk:kk + 1])[0]   # GrottoDcf::EvalAll
                r1 = orc.evalall(p, 1, s0s[kk:kk + 1, 1], cws[kk:kk + 1])[0]
                assert np.array_equal(a0 ^ a1, r0 ^ r1), (n, kk)
            # GrottoDcf::Preprocess + Eval (the lookup, with its e == 0 / e == N rule) on the sampled points
            l0 = orc.grotto_lookup(p, orc.grotto_preprocess(p, 0, s0s[:, 0], cws), xs)
            l1 = orc.grotto_lookup(p, orc.grotto_preprocess(p, 1, s0s[:, 1], cws), xs)
            assert np.array_equal(w0 ^ w1, l0 ^ l1), (n, in_bytes, prg)


def test_tile_chunk_sequence_of_every_scheme_body(emu):
    """The TMA tile accessors (kernels.cuh: CwTileT, CwLmTile) request chunk c + 1 -- or the next tile's chunk 0 -- while the
    scheme body consumes chunk c, so a body must visit the chunk boundaries 0, 1, ..., nchunks - 1 exactly once each and in
    order, for EVERY domain size, or the per-warp pipeline is out of step from the next tile on (the first Grotto walk was:
    its body never touched entry n).  nchunks is what point_kernel computes for the mode."""
    import ctypes
    buf = (ctypes.c_int * 256)()
    for scheme in ("dpf", "dcf", "halftree", "vdpf", "grotto"):
        for n in range(1, 129):
            p = Params(scheme=scheme, in_bits=n, prg="chacha")
            cp, ncw = p.c(), p.ncw
            lpcs = {1}                                   # key-major tiles, modes 4 / 5: two levels per chunk
            if scheme in ("dpf", "halftree", "vdpf"):
                lpcs.add(2)                              # packed rows (mode 6) and level-major tiles (mode 7): four levels
            for lpc in sorted(lpcs):                     # (DCF level-major tiles: two levels = lpc 1, covered above)
                cnt = emu.emul_chunk_sequence(ctypes.byref(cp), lpc, buf, 256)
                want = (ncw + (1 << lpc) - 1) >> lpc
                assert cnt == want and list(buf[:cnt]) == list(range(want)), (scheme, n, lpc, cnt, list(buf[:min(cnt, 8)]))
