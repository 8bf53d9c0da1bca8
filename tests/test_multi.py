"""Multi-device entry points of the C ABI (fssb200_*_multi, one process driving several GPUs).  On a one-GPU box the
shards run as two contexts on the same device (what VERDICT r01 asks for: "shard-vs-full comparison with two contexts on
one device"); with >= 2 GPUs visible they run on cuda:0 and cuda:1.  Checker: the oracle and the unsharded result."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import HASH_KEY_BENCH, Params, synth_inputs


def test_key_shard_matches_python_launcher():
    """fssb200_key_shard (C) == fss_b200.sharding.key_shard (Python): pure arithmetic, no GPU."""
    from fss_b200.multi import key_shard as c_shard
    from fss_b200.sharding import key_shard as py_shard
    for nkeys in (0, 1, 7, 8, 1000, (1 << 22) + 3):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                assert c_shard(nkeys, r, world) == py_shard(nkeys, r, world)
                cover.append(c_shard(nkeys, r, world))
            assert cover[0][0] == 0 and cover[-1][1] == nkeys
            assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    from fss_b200 import _lib as L
    b, e = C.c_size_t(), C.c_size_t()
    assert L.lib.fssb200_key_shard(10, 2, 2, C.byref(b), C.byref(e)) == L.E_INVAL
    assert L.lib.fssb200_eval_multi(None, 1, 0, None, None, None, None, None, None, None, None) == L.E_INVAL


def _devices():
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


def _t(a, d):
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).to(f"cuda:{d}")


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,n,group,prg", [("dpf", 32, "bytes", "aes128_mmo"), ("dcf", 64, "u128", "aes128_mmo"),
                                                 ("halftree", 20, "u64", "chacha")])
def test_eval_multi_matches_oracle(orc, scheme, n, group, prg):
    from fss_b200.multi import MultiContext, key_shard
    devs = _devices()
    p = Params(scheme=scheme, in_bits=n, group=group, prg=prg, hash_key=HASH_KEY_BENCH)
    nkeys = 10001
    s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=n)
    kw = dict(prg=prg, prg_key=p.prg_key, hash_key=p.hash_key)
    mc = MultiContext(devs, scheme, n, group, **kw)
    sh = [key_shard(nkeys, d, len(devs)) for d in range(len(devs))]
    g = mc.gen([_t(s0s[b:e], dv) for (b, e), dv in zip(sh, devs)], [alphas[b:e] for b, e in sh],
               [_t(betas[b:e], dv) for (b, e), dv in zip(sh, devs)])
    cws, ocws = g if scheme == "halftree" else (g, None)
    mc.sync()
    o = orc.gen(p, s0s, alphas, betas, threads=8)
    oc, ooc = o if scheme == "halftree" else (o, None)
    got_cws = np.concatenate([c.cpu().numpy().view(np.uint32) for c in cws])
    from test_gpu_parity import masked
    assert np.array_equal(masked(p, got_cws), masked(p, oc))
    for party in (0, 1):
        ys = mc.eval(party, [_t(s0s[b:e, party], dv) for (b, e), dv in zip(sh, devs)], cws, [xs[b:e] for b, e in sh], ocws)
        mc.sync()
        got = np.concatenate([y.cpu().numpy().view(np.uint32) for y in ys])
        assert np.array_equal(got, orc.eval(p, party, s0s[:, party], oc, xs, ooc, threads=8)), party
    # host arrays of the whole batch, split by the library
    h = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32))  # noqa: E731
    yh = mc.eval_host(1, h(s0s[:, 1]), h(oc), xs, None if ooc is None else h(ooc))
    assert np.array_equal(yh.numpy().view(np.uint32), orc.eval(p, 1, s0s[:, 1], oc, xs, ooc, threads=8))
    mc.close()


@pytest.mark.gpu
def test_eval_all_subtree_shards_equal_the_full_domain(orc):
    """BASELINE configs[3] "subtrees sharded": shard d expands leaf range fssb200_leaf_shard(d) of every key; the
    concatenation must be the unsharded EvalAll (checked against the oracle) -- n = 19, granule 2^17, 4 units."""
    from fss_b200.multi import MultiContext
    devs = _devices()
    p = Params(scheme="dpf", in_bits=19, group="u64")
    s0s, alphas, betas, _ = synth_inputs(p, 3, seed=19)
    oc = orc.gen(p, s0s, alphas, betas)
    want = orc.evalall(p, 0, s0s[:, 0], oc, threads=8)
    mc = MultiContext(devs, "dpf", 19, "u64", prg_key=p.prg_key)
    lr = [mc.leaf_shard(d) for d in range(len(devs))]
    assert lr[0][0] == 0 and sum(c for _, c in lr) == 1 << 19 and lr[1][0] == lr[0][1]
    ys = mc.eval_all(0, [_t(s0s[:, 0], dv) for dv in devs], [_t(oc, dv) for dv in devs], leaf_ranges=lr)
    mc.sync()
    got = np.concatenate([y.cpu().numpy().view(np.uint32) for y in ys], axis=1)
    assert np.array_equal(got, want)
    mc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
def test_eval_host_multi_balanced_split(orc, monkeypatch, pinned):
    """fssb200_eval_host_multi in host mode 1 (or with >= 3 devices): devices claim key blocks from one counter, two calls
    in flight per device (small blocks here so that a test-sized batch is split many times)."""
    from fss_b200.multi import MultiContext
    monkeypatch.setenv("FSSB200_MULTI_MIN_BLOCK_BITS", "12")
    monkeypatch.setenv("FSSB200_MULTI_MAX_BLOCK_BITS", "14")
    ndev = min(max(torch.cuda.device_count(), 2), 4)
    devs = [d % torch.cuda.device_count() for d in range(ndev)]
    p = Params(scheme="dpf", in_bits=32)
    nkeys = 70001
    s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=4)
    oc = orc.gen(p, s0s, alphas, betas, threads=8)
    want = orc.eval(p, 1, s0s[:, 1], oc, xs, threads=8)
    mc = MultiContext(devs, "dpf", 32, "bytes", prg_key=p.prg_key)
    mc.set_host_mode(1)
    h = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32))  # noqa: E731
    a = [h(s0s[:, 1]), h(oc)]
    if pinned:
        a = [t.pin_memory() for t in a]
    launches0 = mc.launch_count()
    yh = mc.eval_host(1, a[0], a[1], xs)
    assert np.array_equal(yh.numpy().view(np.uint32), want)
    assert mc.launch_count() - launches0 >= 5      # several blocks, not one range per device
    mc.close()
