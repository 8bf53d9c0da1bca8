"""VDPF (reference vdpf.cuh, hash/blake3.cuh).

CPU (`-m "not gpu"`): the C restatement against the reference-generated golden fixtures
(tests/golden/golden_vdpf_v1.*, oracle/make_golden_vdpf.py) and, where oracle/_ref exists, against the
compiled reference; the kernels' host-compiled bodies (tests/host_emul) against the oracle; the
properties the reference's own suite checks (src/vdpf_test.cu: reconstruction, proofs of the two
parties agree, a tampered key is rejected).
GPU (`-m gpu`): the CUDA path through the C ABI against the golden fixtures and the oracle.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import Params, Ref, _vp, pack_ints, synth_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
GROUPS = [("bytes", 0), ("u32", 0), ("u64", 0), ("u128", 0), ("u64", 18446744073709551557)]


class VCase:
    def __init__(self, meta, arrays):
        self.meta, self.name = meta, meta["name"]
        self.p = Params(scheme="vdpf", in_bits=meta["in_bits"], group=meta["group"], mod=int(meta["mod"]),
                        prg=meta["prg"], prg_key=bytes.fromhex(meta["prg_key"]), in_bytes=meta["in_bytes"],
                        hash_iv=bytes.fromhex(meta["hash_iv"]), hash=meta.get("hash", "blake3"))
        self.alphas = [int(a) for a in meta["alphas"]]
        self.xs = [int(x) for x in meta["xs"]]
        self._a = arrays

    def __getitem__(self, key):
        return self._a[f"{self.name}/{key}"]


@pytest.fixture(scope="module")
def vgolden():
    arrays, cases = {}, []
    for stem in ("golden_vdpf_v1", "golden_vdpf_sha256_v1"):   # XorHash = Hash = Blake3 / Sha256
        with open(os.path.join(HERE, "golden", stem + ".json")) as f:
            man = json.load(f)
        arrays.update(np.load(os.path.join(HERE, "golden", stem + ".npz")))
        cases += [VCase(m, arrays) for m in man["cases"]]
    return arrays, cases


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- CPU: the checker itself -----------------------------------------------------------------------------------

HASH_KEYS = (("blake3", "hash"), ("sha256", "hash_sha256"))   # plugin, key prefix of its known answers


def test_oracle_hashes_vs_golden(orc, vgolden):
    arrays, _ = vgolden
    for name, pre in HASH_KEYS:
        p = Params(scheme="vdpf", in_bits=8, hash=name)
        assert np.array_equal(orc.hash(p, 0, arrays[pre + "/xor_in"]), arrays[pre + "/xor_out"]), name
        assert np.array_equal(orc.hash(p, 1, arrays[pre + "/hash_in"]), arrays[pre + "/hash_out"]), name


def test_oracle_sha256_vs_hashlib(orc):
    """hash/sha256.cuh:44-89 restated with hashlib: SHA-256(key || msg), and the two digests of (a lsb 0 / 1, b)."""
    p = Params(scheme="vdpf", in_bits=8, hash="sha256")
    key0, key1 = bytes(p.hash_iv)[:16], bytes(p.hash_iv)[32:48]
    rng = np.random.default_rng(5)
    msgs = rng.integers(0, 2 ** 32, size=(40, 4, 4), dtype=np.uint64).astype(np.uint32)
    out = orc.hash(p, 1, msgs)
    ab = rng.integers(0, 2 ** 32, size=(40, 2, 4), dtype=np.uint64).astype(np.uint32)
    outx = orc.hash(p, 0, ab)
    for i in range(40):
        assert out[i].tobytes() == hashlib.sha256(key1 + msgs[i].tobytes()).digest()
        a0, a1 = ab[i, 0].copy(), ab[i, 0].copy()
        a0[3] &= 0xFFFFFFFE
        a1[3] |= 1
        assert outx[i].tobytes() == hashlib.sha256(key0 + a0.tobytes() + ab[i, 1].tobytes()).digest() + \
            hashlib.sha256(key0 + a1.tobytes() + ab[i, 1].tobytes()).digest()


def test_oracle_vs_golden(orc, vgolden):
    _, cases = vgolden
    for c in cases:
        p = c.p
        cws, cs, ocws, status = orc.vdpf_gen(p, c["s0s"], c.alphas, c["betas"])
        assert not status.any()
        assert np.array_equal(cws, c["cws"]) and np.array_equal(cs, c["cs"]) and np.array_equal(ocws, c["ocws"]), c.name
        for party in (0, 1):
            ys, pis = orc.vdpf_eval(p, party, c["s0s"][:, party], cws, cs, ocws, c.xs)
            assert np.array_equal(ys, c[f"ys{party}"]) and np.array_equal(pis, c[f"pis{party}"]), (c.name, party)
            k4 = len(ys) // 4
            if k4:
                assert np.array_equal(orc.vdpf_prove(p, pis[:4 * k4].reshape(k4, 4, 4, 4), cs[:k4]),
                                      c[f"prove{party}"]), c.name
            if c.meta["evalall"] != "none":
                k = c.meta["evalall_keys"]
                ya, pa = orc.vdpf_evalall(p, party, c["s0s"][:k, party], cws[:k], cs[:k], ocws[:k])
                assert np.array_equal(pa, c[f"allpi{party}"]), c.name
                if c.meta["evalall"] == "full":
                    assert np.array_equal(ya, c[f"all{party}"]), c.name
                else:
                    assert [sha(ya[i]) for i in range(k)] == c.meta[f"all{party}_sha256"], c.name


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_compiled_reference(orc):
    ref = Ref()
    for n, hname in [(n, "blake3") for n in (1, 3, 8, 12, 32, 64, 128)] + [(n, "sha256") for n in (1, 8, 12, 32, 64, 128)]:
        for (g, mod) in GROUPS:
            for prg in ("aes128_mmo", "chacha"):
                p = Params(scheme="vdpf", in_bits=n, group=g, mod=mod, prg=prg, hash=hname)
                assert ref.vdpf_supported(p), p
                s0s, alphas, betas, xs = synth_inputs(p, 12, seed=n * 5 + len(g))
                a, b = orc.vdpf_gen(p, s0s, alphas, betas), ref.vdpf_gen(p, s0s, alphas, betas, threads=2)
                for u, v in zip(a, b):
                    assert np.array_equal(u, v), p
                cws, cs, ocws, _ = b
                for party in (0, 1):
                    ra = orc.vdpf_eval(p, party, s0s[:, party], cws, cs, ocws, xs)
                    rb = ref.vdpf_eval(p, party, s0s[:, party], cws, cs, ocws, xs, threads=2)
                    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]), p
                if n <= 8:
                    ra = orc.vdpf_evalall(p, 1, s0s[:2, 1], cws[:2], cs[:2], ocws[:2])
                    rb = ref.vdpf_evalall(p, 1, s0s[:2, 1], cws[:2], cs[:2], ocws[:2])
                    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]), p


def test_vdpf_properties(orc):
    """src/vdpf_test.cu: shares reconstruct the point function, honest proofs verify, a flipped
    correction word is rejected."""
    for (g, mod), hname in zip(GROUPS * 2, ["blake3"] * 5 + ["sha256", ("sha256", "blake3"), ("blake3", "sha256"), "sha256", "sha256"]):
        p = Params(scheme="vdpf", in_bits=10, group=g, mod=mod, hash=hname)
        s0s, alphas, betas, _ = synth_inputs(p, 3, seed=9)
        cws, cs, ocws, status = orc.vdpf_gen(p, s0s, alphas, betas)
        assert not status.any()
        y0, pi0 = orc.vdpf_evalall(p, 0, s0s[:, 0], cws, cs, ocws)
        y1, pi1 = orc.vdpf_evalall(p, 1, s0s[:, 1], cws, cs, ocws)
        assert np.array_equal(pi0, pi1)                                  # Verify accepts
        pd = Params(scheme="dpf", in_bits=10, group=g, mod=mod)           # group helper only
        rec = orc.group_add(pd, y0.reshape(-1, 4), y1.reshape(-1, 4)).reshape(y0.shape)
        want = np.zeros_like(rec)
        beta_g = orc.group_add(pd, betas, np.zeros_like(betas))           # From/Into normalisation of beta
        for k in range(3):
            want[k, alphas[k]] = beta_g[k]
        assert np.array_equal(rec, want), g
        bad = cws.copy()
        bad[0, 4, 0] ^= 0x10                                              # party 1 receives a tampered key
        _, pib = orc.vdpf_evalall(p, 1, s0s[:, 1], bad, cs, ocws)
        assert not np.array_equal(pi0[0], pib[0]) and np.array_equal(pi0[1:], pib[1:])


# ---- CPU: kernel bodies compiled for the host ------------------------------------------------------------------------

@pytest.fixture(scope="module")
def emu():
    from test_host_emul import emu as _emu_fixture  # reuse the build logic
    return _emu_fixture.__wrapped__()


def test_kernel_bodies_match_oracle(emu, orc, vgolden):
    arrays, _ = vgolden
    for hname, pre in HASH_KEYS:
        cp = Params(scheme="vdpf", in_bits=8, hash=hname).c()
        for which, key_in, key_out, shape in ((0, "/xor_in", "/xor_out", (4, 4)), (1, "/hash_in", "/hash_out", (2, 4))):
            out = np.zeros((len(arrays[pre + key_in]),) + shape, np.uint32)
            emu.emul_hash(C.byref(cp), which, C.c_size_t(len(out)), _vp(np.ascontiguousarray(arrays[pre + key_in])), _vp(out))
            assert np.array_equal(out, arrays[pre + key_out]), hname
    for n in (1, 2, 5, 8, 12, 32, 33, 64, 65, 128):
        for gi, (g, mod) in enumerate(GROUPS):
            for prg in ("aes128_mmo", "chacha"):
                hname = ("blake3", "sha256", ("sha256", "blake3"), ("blake3", "sha256"))[(n + gi) % 4]
                p = Params(scheme="vdpf", in_bits=n, group=g, mod=mod, prg=prg, hash=hname)
                k = 40
                s0s, alphas, betas, xs = synth_inputs(p, k, seed=n + len(g))
                xs[1], xs[2] = 0, (1 << n) - 1
                want = orc.vdpf_gen(p, s0s, alphas, betas)
                cws, cs = np.zeros((k, n, 8), np.uint32), np.zeros((k, 4, 4), np.uint32)
                ocws, status = np.zeros((k, 4), np.uint32), np.zeros(k, np.int32)
                cp, al = p.c(), pack_ints(alphas, p.in_bytes)
                emu.emul_vdpf_gen(C.byref(cp), C.c_size_t(k), _vp(np.ascontiguousarray(s0s)), _vp(al), _vp(betas),
                                  _vp(cws), _vp(cs), _vp(ocws), _vp(status))
                for u, v in zip(want, (cws, cs, ocws, status)):
                    assert np.array_equal(u, v), (n, g, prg)
                xb = pack_ints(xs, p.in_bytes)
                pd = Params(scheme="dpf", in_bits=n, group=g, mod=mod, prg=prg)
                cw_s, _, extra, _ = orc.relayout(pd, np.concatenate([cws, np.zeros((k, 1, 8), np.uint32)], axis=1))
                for party in (0, 1):
                    wy, wp = orc.vdpf_eval(p, party, s0s[:, party], cws, cs, ocws, xs)
                    for lm in (0, 1):
                        ys, pis = np.zeros((k, 4), np.uint32), np.zeros((k, 4, 4), np.uint32)
                        emu.emul_vdpf_eval(C.byref(cp), party, C.c_size_t(k), _vp(np.ascontiguousarray(s0s[:, party])),
                                           _vp(cws), _vp(cs), _vp(ocws), _vp(xb), _vp(ys), _vp(pis), lm,
                                           _vp(np.ascontiguousarray(cw_s)), _vp(np.ascontiguousarray(extra)))
                        assert np.array_equal(ys, wy) and np.array_equal(pis, wp), (n, g, prg, party, lm)
                pts = np.ascontiguousarray(wp.reshape(10, 4, 4, 4))
                got = np.zeros((10, 4, 4), np.uint32)
                emu.emul_vdpf_prove(C.byref(cp), C.c_size_t(10), C.c_size_t(4), _vp(pts), _vp(cs[:10].copy()), _vp(got))
                assert np.array_equal(got, orc.vdpf_prove(p, pts, cs[:10]))


# ---- GPU: the CUDA path through the C ABI ------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _t(a, dev):
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(dev)


def _n(t):
    return t.detach().cpu().contiguous().numpy().view(np.uint32)


def _ctx(p):
    import fss_b200
    return fss_b200.Context("vdpf", p.in_bits, p.group, mod=p.mod, prg=p.prg, prg_key=p.prg_key, in_bytes=p.in_bytes,
                            hash_iv=bytes(p.hash_iv), hash=p.hash)


@pytest.mark.gpu
def test_gpu_hashes(dev, vgolden):
    arrays, _ = vgolden
    for hname, pre in HASH_KEYS:
        ctx = _ctx(Params(scheme="vdpf", in_bits=8, hash=hname))
        assert np.array_equal(_n(ctx.hash(0, _t(arrays[pre + "/xor_in"], dev))), arrays[pre + "/xor_out"]), hname
        assert np.array_equal(_n(ctx.hash(1, _t(arrays[pre + "/hash_in"], dev))), arrays[pre + "/hash_out"]), hname


@pytest.mark.gpu
def test_gpu_golden(dev, vgolden):
    _, cases = vgolden
    for c in cases:
        p, ctx = c.p, _ctx(c.p)
        cws, cs, ocws, status = ctx.vdpf_gen(_t(c["s0s"], dev), c.alphas, _t(c["betas"], dev))
        assert not _n(status).any()
        assert np.array_equal(_n(cws), c["cws"]) and np.array_equal(_n(cs), c["cs"]), c.name
        assert np.array_equal(_n(ocws), c["ocws"]), c.name
        for party in (0, 1):
            ys, pis = ctx.vdpf_eval(party, _t(c["s0s"][:, party], dev), cws, cs, ocws, c.xs)
            assert np.array_equal(_n(ys), c[f"ys{party}"]) and np.array_equal(_n(pis), c[f"pis{party}"]), (c.name, party)
            k4 = ys.shape[0] // 4
            if k4:
                pr = ctx.vdpf_prove(pis[:4 * k4].reshape(k4, 4, 4, 4), cs[:k4])
                assert np.array_equal(_n(pr), c[f"prove{party}"]), c.name
            if c.meta["evalall"] != "none":
                k = c.meta["evalall_keys"]
                ya, pa = ctx.vdpf_eval_all(party, _t(c["s0s"][:k, party], dev), cws[:k], cs[:k], ocws[:k])
                assert np.array_equal(_n(pa), c[f"allpi{party}"]), c.name
                if c.meta["evalall"] == "full":
                    assert np.array_equal(_n(ya), c[f"all{party}"]), c.name
                else:
                    assert [sha(_n(ya)[i]) for i in range(k)] == c.meta[f"all{party}_sha256"], c.name


@pytest.mark.gpu
@pytest.mark.parametrize("hname", ["blake3", "sha256", ("sha256", "blake3"), ("blake3", "sha256")])
@pytest.mark.parametrize("n,group,mod,prg,nkeys", [
    (32, "bytes", 0, "aes128_mmo", 3000), (32, "u64", 0, "chacha", 1500), (64, "u128", 0, "aes128_mmo", 1031),
    (20, "u32", 0, "aes128_mmo", 777), (128, "u64", 18446744073709551557, "aes128_mmo", 257), (1, "bytes", 0, "chacha", 33),
    (33, "u64", 0, "aes128_mmo", 100),
])
def test_gpu_random_batches(dev, orc, n, group, mod, prg, nkeys, hname):
    import torch
    if not isinstance(hname, str) and n not in (32, 1):
        pytest.skip("mixed hash pairs: two shapes are enough")
    p = Params(scheme="vdpf", in_bits=n, group=group, mod=mod, prg=prg, hash=hname)
    ctx = _ctx(p)
    s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=n + nkeys)
    xs[1], xs[2] = 0, (1 << n) - 1
    want = orc.vdpf_gen(p, s0s, alphas, betas, threads=8)
    got = ctx.vdpf_gen(_t(s0s, dev), alphas, _t(betas, dev))
    for u, v in zip(want, got):
        assert np.array_equal(u, _n(v).view(u.dtype).reshape(u.shape))
    cws, cs, ocws, _ = got
    lay = ctx.relayout(cws)
    pis = []
    for party in (0, 1):
        wy, wp = orc.vdpf_eval(p, party, s0s[:, party], want[0], want[1], want[2], xs, threads=8)
        ys, pt = ctx.vdpf_eval(party, _t(s0s[:, party], dev), cws, cs, ocws, xs)
        assert np.array_equal(_n(ys), wy) and np.array_equal(_n(pt), wp), party
        ys2, pt2 = ctx.vdpf_eval(party, _t(s0s[:, party], dev), cws, cs, ocws, xs, layout=(lay[0], lay[2]))
        assert torch.equal(ys, ys2) and torch.equal(pt, pt2)
        pis.append(pt)
    assert torch.equal(pis[0], pis[1]) and bool(ctx.vdpf_verify(pis[0], pis[1]).all())
    # host-buffer entry points, chunks smaller than the batch
    ctx.reserve_host(max(1, nkeys // 3))
    hg = ctx.vdpf_gen(_t(s0s, "cpu"), alphas, _t(betas, "cpu"))
    for u, v in zip(want, hg):
        assert v.device.type == "cpu" and np.array_equal(u, _n(v).view(u.dtype).reshape(u.shape))
    hy, hp = ctx.vdpf_eval(1, _t(s0s[:, 1], "cpu"), hg[0], hg[1], hg[2], xs)
    assert np.array_equal(_n(hy), wy) and np.array_equal(_n(hp), wp)


@pytest.mark.gpu
@pytest.mark.parametrize("n,group,prg,nkeys", [(10, "u64", "aes128_mmo", 5), (3, "bytes", "chacha", 9),
                                              (14, "u128", "aes128_mmo", 2), (18, "bytes", "aes128_mmo", 1)])
def test_gpu_evalall_and_tamper(dev, orc, n, group, prg, nkeys):
    import torch
    p = Params(scheme="vdpf", in_bits=n, group=group, prg=prg, hash="sha256" if n in (3, 14) else "blake3")
    ctx = _ctx(p)
    s0s, alphas, betas, _ = synth_inputs(p, nkeys, seed=n)
    cws, cs, ocws, status = ctx.vdpf_gen(_t(s0s, dev), alphas, _t(betas, dev))
    res = []
    for party in (0, 1):
        ya, pa = ctx.vdpf_eval_all(party, _t(s0s[:, party], dev), cws, cs, ocws)
        wy, wp = orc.vdpf_evalall(p, party, s0s[:, party], _n(cws), _n(cs), _n(ocws), threads=4)
        assert np.array_equal(_n(ya), wy) and np.array_equal(_n(pa), wp), party
        res.append((ya, pa))
    assert bool(ctx.vdpf_verify(res[0][1], res[1][1]).all())
    bad = cws.clone()
    bad[0, n // 2, 1] ^= 4
    _, pb = ctx.vdpf_eval_all(1, _t(s0s[:, 1], dev), bad, cs, ocws)
    ok = ctx.vdpf_verify(res[0][1], pb)
    assert not bool(ok[0]) and bool(ok[1:].all())


@pytest.mark.gpu
def test_gpu_vdpf_scheme_errors(dev):
    import torch
    import fss_b200
    from fss_b200 import _lib as L
    ctx = fss_b200.Context("vdpf", 16)
    z = torch.zeros((4, 17, 8), dtype=torch.int32, device=dev)
    s = torch.zeros((4, 4), dtype=torch.int32, device=dev)
    with pytest.raises(L.FssError) as e:
        ctx.eval(0, s, z, [1, 2, 3, 4])                # generic entry points do not apply to a VDPF context
    assert e.value.code == L.E_SCHEME
    dpf = fss_b200.Context("dpf", 16)
    with pytest.raises(L.FssError) as e:
        dpf.vdpf_eval(0, s, z[:, :16], torch.zeros((4, 4, 4), dtype=torch.int32, device=dev), s, [1, 2, 3, 4])
    assert e.value.code == L.E_SCHEME


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
def test_gpu_evalall_lanes_per_key(dev, orc, monkeypatch, lanes):
    """vdpf_finish_kernel<G, LPK>: every lanes-per-key variant (the launcher picks one from the key count) against the
    oracle -- ragged key groups in the last warp, domains smaller and larger than LPK, both parties."""
    monkeypatch.setenv("FSSB200_VDPF_FINISH_LANES", str(lanes))
    for n, group, nkeys in ((6, "u64", 37), (3, "bytes", 5), (1, "u32", 67), (9, "u128", 130)):
        p = Params(scheme="vdpf", in_bits=n, group=group, hash="sha256" if n in (6, 1) else "blake3")
        ctx = _ctx(p)
        s0s, alphas, betas, _ = synth_inputs(p, nkeys, seed=lanes * 100 + n)
        cws, cs, ocws, _ = orc.vdpf_gen(p, s0s, alphas, betas, threads=4)
        for party in (0, 1):
            ya, pa = ctx.vdpf_eval_all(party, _t(s0s[:, party], dev), _t(cws, dev), _t(cs, dev), _t(ocws, dev))
            wy, wp = orc.vdpf_evalall(p, party, s0s[:, party], cws, cs, ocws, threads=4)
            assert np.array_equal(_n(ya), wy) and np.array_equal(_n(pa), wp), (n, group, party)
