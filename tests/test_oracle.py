"""CPU tests of the checker itself: the C restatement (oracle/fss_oracle.c) against

* the literal known-answer vectors the survey took from the reference,
* every golden fixture generated from the compiled reference (tests/golden/, oracle/make_golden.py),
* the compiled reference itself where oracle/_ref is available (this container),
* the reconstruction properties the reference's own GTest suites check (src/*_test.cu).
"""
import hashlib

import numpy as np
import pytest

from oracle import AES_KEYS, CHACHA_NONCE, HASH_KEY_BENCH, HASH_KEY_SAMPLE, Params, synth_inputs

FIX_SEEDS = np.array([[[0x11111111, 0x22222222, 0x33333333, 0x44444440],
                       [0x55555555, 0x66666666, 0x77777777, 0x88888880]]], dtype=np.uint32)
FIX_BETA = np.array([[7, 0, 0, 0]], dtype=np.uint32)


def h(s):
    return np.array([int(w, 16) for w in s.split()], dtype=np.uint32)


# ---- SURVEY.md section 8c literal vectors ---------------------------------------------------------------

def test_kat_prg(orc):
    out = orc.prg_gen(Params(prg="aes128_mmo"), 4, FIX_SEEDS[0, :1])[0]
    assert np.array_equal(out[0], h("1423e6d2 60533602 813b3fbc 412b31dc"))
    assert np.array_equal(out[1], h("2d18cbe7 a5eddcc9 c691e4f6 97831704"))
    assert np.array_equal(out[2], h("2c046c24 8a033811 cb0335b6 3d799ddb"))
    assert np.array_equal(out[3], h("2413b69f ffed7ac4 314804c9 595d2580"))
    out = orc.prg_gen(Params(prg="chacha"), 2, FIX_SEEDS[0, :1])[0]
    assert np.array_equal(out[0], h("78a77dcf c5f06068 728fd3bc 0cc9707f"))
    assert np.array_equal(out[1], h("92b1a797 953bfae1 da7ec472 b822db04"))


def test_kat_dpf_n8(orc):
    p = Params(scheme="dpf", in_bits=8, in_bytes=1)
    cws = orc.gen(p, FIX_SEEDS, [42], FIX_BETA)
    assert np.array_equal(cws[0, 0, :4], h("84863059 b57c0062 c0a8015f 3520c9e8"))
    assert cws[0, 0, 4] & 0xFF == 1
    assert np.array_equal(cws[0, 8, :4], h("0b17cf59 cd54e225 ef2cd79f 30ad9cf8"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [42])[0], h("a81892ad ad2e583e 272d1483 d2bfa094"))
    assert np.array_equal(orc.eval(p, 1, FIX_SEEDS[:, 1], cws, [42])[0], h("a81892aa ad2e583e 272d1483 d2bfa094"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [100])[0], h("7e2be32f 4508eb0c 537f8e89 39b1bd96"))


def test_kat_dpf_n32(orc):
    p = Params(scheme="dpf", in_bits=32)
    cws = orc.gen(p, FIX_SEEDS, [42], FIX_BETA)
    assert np.array_equal(cws[0, 32, :4], h("03627e21 c3ebef2b 2d3d7855 909643de"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [42])[0], h("2dde1ed3 9b45aee5 f8f2e367 485e080c"))
    assert np.array_equal(orc.eval(p, 1, FIX_SEEDS[:, 1], cws, [42])[0], h("2dde1ed4 9b45aee5 f8f2e367 485e080c"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [0xDEADBEEF])[0],
                          h("2b6fd50b 54d87a3e 26f3ee9b 01c8dd7e"))


def test_kat_dcf(orc):
    p = Params(scheme="dcf", in_bits=8, in_bytes=1)
    cws = orc.gen(p, FIX_SEEDS, [42], FIX_BETA)
    assert np.array_equal(cws[0, 0, :4], h("79f28305 7b1a2718 c9874166 fe3afd02"))
    assert np.array_equal(cws[0, 0, 4:], h("0b706acf e72e35b7 1984be09 3d5a09d1"))
    assert np.array_equal(cws[0, 8, 4:], h("7239d2cd ca3a5355 a277702d 22f5efca"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [10])[0], h("3a28e8cb 5f11990b f2fee90f a7604478"))
    assert np.array_equal(orc.eval(p, 1, FIX_SEEDS[:, 1], cws, [10])[0], h("3a28e8cc 5f11990b f2fee90f a7604478"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [200])[0], h("e386e89c fdd3d119 1691e062 90380280"))
    p = Params(scheme="dcf", in_bits=64, group="u128")
    cws = orc.gen(p, FIX_SEEDS, [42], FIX_BETA)
    assert np.array_equal(cws[0, 64, 4:], h("6675d222 a63e84c9 d6f446d0 985772b0"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [10])[0], h("865812cd e91b6533 fa4486b0 9bc4ab6c"))
    assert np.array_equal(orc.eval(p, 1, FIX_SEEDS[:, 1], cws, [10])[0], h("79a7ed3a 16e49acc 05bb794f 643b5492"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [1 << 40])[0],
                          h("25494a62 f745e3eb 2e8feeb2 48222122"))


def test_kat_halftree_grotto(orc):
    p = Params(scheme="halftree", in_bits=8, in_bytes=1, hash_key=HASH_KEY_SAMPLE)
    cws, ocws = orc.gen(p, FIX_SEEDS, [42], FIX_BETA)
    assert np.array_equal(ocws[0], h("b3205fb1 a28dd128 8cc01d96 da93f6e8"))
    assert np.array_equal(cws[0, 7, :4], h("282225a6 ff32b373 aee8bd0c 83bb9b71"))
    assert cws[0, 7, 4] & 0xFF == 0
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [42], ocws)[0],
                          h("e58e28f0 dfccd22d 550c9b5d 12125964"))
    assert np.array_equal(orc.eval(p, 1, FIX_SEEDS[:, 1], cws, [42], ocws)[0],
                          h("e58e28f7 dfccd22d 550c9b5d 12125964"))
    assert np.array_equal(orc.eval(p, 0, FIX_SEEDS[:, 0], cws, [100], ocws)[0],
                          h("4d029550 d8e97069 b09bc4aa 052f753a"))
    p = Params(scheme="grotto", in_bits=8, in_bytes=1)
    cws = orc.gen(p, FIX_SEEDS, [42])
    ys = orc.evalall(p, 0, FIX_SEEDS[:, 0], cws)[0]
    assert "".join(map(str, ys[:16])) == "1011111110111000"
    assert "".join(map(str, ys[40:48])) == "11111110"


# ---- golden fixtures generated from the compiled reference -------------------------------------------------

def test_golden_prg(orc, golden):
    seeds = golden.arrays["prg/seeds"]
    for prg in ("aes128_mmo", "chacha"):
        for mul in (1, 2, 4):
            assert np.array_equal(orc.prg_gen(Params(prg=prg), mul, seeds), golden.arrays[f"prg/{prg}_{mul}"])


def test_golden_cases(orc, golden):
    assert len(golden.cases) >= 100
    for c in golden.cases:
        p, s0s = c.p, c["s0s"]
        r = orc.gen(p, s0s, c.alphas, c.betas)
        cws, ocws = r if p.scheme == "halftree" else (r, None)
        assert np.array_equal(c.masked_cws(cws), c.masked_cws(c["cws"])), c.name
        if ocws is not None:
            assert np.array_equal(ocws, c["ocws"]), c.name
        for party in (0, 1):
            if p.scheme != "grotto":
                ys = orc.eval(p, party, s0s[:, party], c["cws"], c.xs, ocws)
                assert np.array_equal(ys, c[f"ys{party}"]), (c.name, party)
            mode = c.meta["evalall"]
            if mode == "none":
                continue
            k = c.meta["evalall_keys"]
            if p.in_bits > 24:
                continue  # 2^28 leaves take ~1 min per party in the plain-C oracle; covered on the GPU
            ya = orc.evalall(p, party, s0s[:k, party], c["cws"][:k], None if ocws is None else ocws[:k])
            if mode == "full":
                assert np.array_equal(ya, c[f"all{party}"]), (c.name, party)
            else:
                digests = [hashlib.sha256(np.ascontiguousarray(ya[i]).tobytes()).hexdigest() for i in range(k)]
                assert digests == c.meta[f"all{party}_sha256"], (c.name, party)
            if c.has(f"pt{party}"):
                pt = orc.grotto_preprocess(p, party, s0s[:k, party], c["cws"][:k])
                assert np.array_equal(pt, c[f"pt{party}"]), c.name
                assert np.array_equal(orc.grotto_lookup(p, pt, c.xs[:k]), c[f"lookup{party}"]), c.name


# ---- against the compiled reference, fresh random inputs ----------------------------------------------------

@pytest.mark.parametrize("prg", ["aes128_mmo", "chacha"])
def test_against_compiled_reference(orc, ref, prg):
    cfgs = [("dpf", 32, "bytes", 0), ("dcf", 64, "u128", 1 << 127), ("halftree", 32, "bytes", 0),
            ("dcf", 32, "u64", 18446744073709551557), ("dpf", 16, "u8", 251), ("dpf", 128, "u128", 1 << 127),
            ("dcf", 100, "u64", 0), ("halftree", 1, "u64", 0), ("dpf", 1, "bytes", 0)]
    for scheme, n, group, mod in cfgs:
        p = Params(scheme=scheme, in_bits=n, group=group, mod=mod, prg=prg, hash_key=HASH_KEY_BENCH)
        assert ref.supported(p), p
        s0s, alphas, betas, xs = synth_inputs(p, 64, seed=1000 + n)
        r, o = ref.gen(p, s0s, alphas, betas), orc.gen(p, s0s, alphas, betas)
        rc, roc = r if scheme == "halftree" else (r, None)
        oc, ooc = o if scheme == "halftree" else (o, None)
        from golden_util import Case
        fake = Case.__new__(Case)
        fake.p = p
        assert np.array_equal(fake.masked_cws(rc), fake.masked_cws(oc))
        if roc is not None:
            assert np.array_equal(roc, ooc)
        for party in (0, 1):
            assert np.array_equal(ref.eval(p, party, s0s[:, party], rc, xs, roc),
                                  orc.eval(p, party, s0s[:, party], rc, xs, roc))


def test_evalall_against_compiled_reference(orc, ref):
    for scheme, n, group in [("dpf", 12, "bytes"), ("dcf", 12, "u64"), ("halftree", 12, "u64"),
                             ("grotto", 12, "bytes")]:
        p = Params(scheme=scheme, in_bits=n, group=group, hash_key=HASH_KEY_BENCH)
        s0s, alphas, betas, xs = synth_inputs(p, 3, seed=77)
        r = ref.gen(p, s0s, alphas, None if scheme == "grotto" else betas)
        cws, ocws = r if scheme == "halftree" else (r, None)
        for party in (0, 1):
            assert np.array_equal(ref.evalall(p, party, s0s[:, party], cws, ocws),
                                  orc.evalall(p, party, s0s[:, party], cws, ocws))


# ---- properties the reference's own tests check (src/dpf_test.cu:58-139, dcf_test.cu:58-135, ...) ---------------

@pytest.mark.parametrize("group,mod", [("bytes", 0), ("u64", 0), ("u128", 1 << 127), ("u32", 4294967291)])
@pytest.mark.parametrize("scheme", ["dpf", "dcf", "halftree"])
def test_reconstruction(orc, scheme, group, mod):
    n = 16
    p = Params(scheme=scheme, in_bits=n, group=group, mod=mod, hash_key=HASH_KEY_BENCH)
    k = 6
    s0s, alphas, betas, _ = synth_inputs(p, k, seed=5)
    alphas[0], alphas[1] = 107, (1 << n) - 1
    if group != "bytes":  # beta must be a canonical group element for the equality below
        betas = orc.group_add(p, betas, np.zeros_like(betas))
    r = orc.gen(p, s0s, alphas, betas)
    cws, ocws = r if scheme == "halftree" else (r, None)
    y0 = orc.evalall(p, 0, s0s[:, 0], cws, ocws)
    y1 = orc.evalall(p, 1, s0s[:, 1], cws, ocws)
    tot = orc.group_add(p, y0.reshape(-1, 4), y1.reshape(-1, 4)).reshape(k, 1 << n, 4)
    for i in range(k):
        expect = np.zeros((1 << n, 4), dtype=np.uint32)
        if scheme == "dcf":
            expect[: alphas[i]] = betas[i]          # y = beta for x < alpha (kLt)
        else:
            expect[alphas[i]] = betas[i]
        assert np.array_equal(tot[i], expect), (scheme, group, i)
    # EvalAll == Eval on every sampled point
    xs = [0, 1, 107, 108, (1 << n) - 1, 12345]
    for x in xs:
        ys = orc.eval(p, 0, s0s[:, 0], cws, [x] * k, ocws)
        assert np.array_equal(ys, y0[:, x])


def test_grotto_properties(orc):
    n = 10
    p = Params(scheme="grotto", in_bits=n)
    s0s, alphas, _, xs = synth_inputs(p, 5, seed=9)
    alphas[0], alphas[1] = 0, (1 << n) - 1       # edge cases of src/grotto_dcf_test.cu:99-137
    cws = orc.gen(p, s0s, alphas)
    y0, y1 = orc.evalall(p, 0, s0s[:, 0], cws), orc.evalall(p, 1, s0s[:, 1], cws)
    for i in range(5):
        want = (np.arange(1 << n) >= alphas[i]).astype(np.uint8)
        assert np.array_equal(y0[i] ^ y1[i], want)
    for party, ya in ((0, y0), (1, y1)):
        pt = orc.grotto_preprocess(p, party, s0s[:, party], cws)
        for x in (0, 1, 511, 1022, 1023):
            assert np.array_equal(orc.grotto_lookup(p, pt, [x] * 5), ya[:, x])  # grotto_dcf_test.cu:79-96


def test_leaf_ranges(orc):
    for scheme, group in (("dpf", "bytes"), ("dcf", "u64"), ("halftree", "u128")):
        p = Params(scheme=scheme, in_bits=9, group=group, hash_key=HASH_KEY_BENCH)
        s0s, alphas, betas, _ = synth_inputs(p, 2, seed=3)
        r = orc.gen(p, s0s, alphas, betas)
        cws, ocws = r if scheme == "halftree" else (r, None)
        full = orc.evalall(p, 1, s0s[:, 1], cws, ocws)
        for b, cnt in ((0, 64), (64, 64), (128, 384), (511, 1), (3, 5)):
            assert np.array_equal(orc.evalall(p, 1, s0s[:, 1], cws, ocws, b, cnt), full[:, b:b + cnt])
