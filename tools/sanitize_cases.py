#!/usr/bin/env python
"""Small invocation of every kernel family, checked against the oracle -- the workload that
`tools/gpu_session.sh sanitize` runs under compute-sanitizer (memcheck / racecheck / synccheck / initcheck).
Sizes are ragged on purpose (tiles of 32 keys, 768-thread CTAs, TMA boxes that run off the tensor)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fss_b200  # noqa: E402
from oracle import Orc, Params, synth_inputs  # noqa: E402

dev = torch.device("cuda:0")
orc = Orc()
only = set(sys.argv[1:])


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)


def same(name, got, want):
    got = got.cpu().numpy().view(want.dtype).reshape(want.shape)
    assert np.array_equal(got, want), name
    print("ok", name, flush=True)


for scheme, n, group, prg, nkeys in (("dpf", 32, "bytes", "aes128_mmo", 1000), ("dpf", 17, "u64", "chacha", 333),
                                     ("dcf", 64, "u128", "aes128_mmo", 777), ("dcf", 9, "u32", "chacha", 100),
                                     ("halftree", 32, "bytes", "aes128_mmo", 901), ("halftree", 5, "bytes", "chacha", 65)):
    tag = f"{scheme}-{n}-{group}-{prg}"
    if only and "point" not in only:
        break
    p = Params(scheme=scheme, in_bits=n, group=group, prg=prg)
    s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=n)
    ctx = fss_b200.Context(scheme, n, group, p.mod, prg, p.pred, p.prg_key, p.hash_key, p.in_bytes)
    r = ctx.gen(t(s0s), alphas, t(betas))
    cws, ocws = r if scheme == "halftree" else (r, None)
    cw_np = cws.cpu().numpy().view(np.uint32)
    oc_np = None if ocws is None else ocws.cpu().numpy().view(np.uint32)
    for party in (0, 1):
        ys = ctx.eval(party, t(s0s[:, party]), cws, xs, ocws=ocws)
        w = orc.eval(p, party, s0s[:, party], cw_np, xs, oc_np)
        same(f"eval {tag} party {party}", ys, w)
    lay = ctx.relayout(cws)
    ys2 = ctx.eval_levelmajor(0, t(s0s[:, 0]), lay, xs, ocws=ocws)
    same(f"levelmajor {tag}", ys2, orc.eval(p, 0, s0s[:, 0], cw_np, xs, oc_np))

if not only or "evalall" in only:
    for scheme, n, group, prg, nkeys in (("dpf", 18, "bytes", "aes128_mmo", 2), ("dpf", 11, "u64", "chacha", 3),
                                         ("halftree", 18, "bytes", "aes128_mmo", 2), ("dcf", 17, "u128", "aes128_mmo", 1),
                                         ("grotto", 18, "bytes", "aes128_mmo", 2),
                                         # cooperative bottom stage with the shortest walks / both PRGs
                                         ("dpf", 13, "u32", "aes128_mmo", 3), ("halftree", 12, "bytes", "chacha", 2),
                                         ("dpf", 14, "u128", "chacha", 1)):
        tag = f"{scheme}-{n}-{group}-{prg}"
        p = Params(scheme=scheme, in_bits=n, group=group, prg=prg)
        s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=100 + n)
        ctx = fss_b200.Context(scheme, n, group, p.mod, prg, p.pred, p.prg_key, p.hash_key, p.in_bytes)
        r = ctx.gen(t(s0s), alphas, None if scheme == "grotto" else t(betas))
        cws, ocws = r if scheme == "halftree" else (r, None)
        cw_np = cws.cpu().numpy().view(np.uint32)
        oc_np = None if ocws is None else ocws.cpu().numpy().view(np.uint32)
        ya = ctx.eval_all(1, t(s0s[:, 1]), cws, ocws=ocws)
        w = orc.evalall(p, 1, s0s[:, 1], cw_np, oc_np)
        same(f"eval_all {tag}", ya, w)
        if scheme == "grotto":
            pt = ctx.grotto_preprocess(0, t(s0s[:, 0]), cws)
            yl = ctx.grotto_lookup(pt, xs)
            w0 = orc.evalall(p, 0, s0s[:, 0], cw_np)
            same(f"grotto lookup {tag}", yl, np.array([w0[k][int(x)] for k, x in enumerate(xs)], dtype=np.uint8))

if not only or "vdpf" in only:
    ctx = fss_b200.Context("vdpf", 16, "bytes", prg="aes128_mmo")
    g = torch.Generator(device=dev).manual_seed(5)
    nk = 500
    s0s = torch.randint(-2 ** 31, 2 ** 31, (nk, 2, 4), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    betas = torch.randint(-2 ** 31, 2 ** 31, (nk, 4), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    s0s[:, :, 3] &= ~1
    betas[:, 3] &= ~1
    alphas = torch.randint(0, 1 << 16, (nk,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    cws, cs, ocws, status = ctx.vdpf_gen(s0s, alphas, betas)
    ok = status == 0
    y0, p0 = ctx.vdpf_eval(0, s0s[:, 0].contiguous(), cws, cs, ocws, alphas)
    y1, p1 = ctx.vdpf_eval(1, s0s[:, 1].contiguous(), cws, cs, ocws, alphas)
    assert torch.equal((y0 ^ y1)[ok], betas[ok]) and torch.equal(p0[ok], p1[ok])
    ya0, pa0 = ctx.vdpf_eval_all(0, s0s[:4, 0].contiguous(), cws[:4], cs[:4], ocws[:4])
    ya1, pa1 = ctx.vdpf_eval_all(1, s0s[:4, 1].contiguous(), cws[:4], cs[:4], ocws[:4])
    assert torch.equal(pa0, pa1)
    print("ok vdpf", flush=True)

if not only or "vdpf2" in only:
    # SHA-256 hash plugins and a narrow lanes-per-key variant of the finish kernel (ragged key groups in the last warp)
    os.environ["FSSB200_VDPF_FINISH_LANES"] = "4"
    for hname, n, nk in (("sha256", 6, 37), (("blake3", "sha256"), 9, 130)):
        p = Params(scheme="vdpf", in_bits=n, group="u64", hash=hname)
        ctx = fss_b200.Context("vdpf", n, "u64", prg_key=p.prg_key, hash_iv=bytes(p.hash_iv), hash=hname)
        s0s, alphas, betas, xs = synth_inputs(p, nk, seed=n)
        want = orc.vdpf_gen(p, s0s, alphas, betas, threads=4)
        got = ctx.vdpf_gen(t(s0s), alphas, t(betas))
        for u, v in zip(want[:3], got[:3]):
            same(f"vdpf gen {hname}", v, u)
        wy, wp = orc.vdpf_eval(p, 1, s0s[:, 1], want[0], want[1], want[2], xs, threads=4)
        ys, pis = ctx.vdpf_eval(1, t(s0s[:, 1]), got[0], got[1], got[2], xs)
        same(f"vdpf eval {hname}", ys, wy)
        same(f"vdpf eval proofs {hname}", pis, wp)
        way, wpa = orc.vdpf_evalall(p, 0, s0s[:, 0], want[0], want[1], want[2], threads=4)
        ya, pa = ctx.vdpf_eval_all(0, t(s0s[:, 0]), got[0], got[1], got[2])
        same(f"vdpf evalall {hname}", ya, way)
        same(f"vdpf evalall proofs {hname}", pa, wpa)
    os.environ.pop("FSSB200_VDPF_FINISH_LANES")

if not only or "host" in only:
    p = Params(scheme="dpf", in_bits=32)
    s0s, alphas, betas, xs = synth_inputs(p, 3000, seed=9)
    ctx = fss_b200.Context("dpf", 32, "bytes", prg="aes128_mmo")
    ctx.reserve_host(1024)
    hs = torch.from_numpy(s0s.view(np.int32)).pin_memory()
    cws = ctx.gen(hs, alphas, torch.from_numpy(betas.view(np.int32)).pin_memory())
    ys = ctx.eval(0, hs[:, 0].contiguous().pin_memory(), cws.pin_memory(), xs)
    same("eval_host", ys, orc.eval(p, 0, s0s[:, 0], cws.numpy().view(np.uint32), xs))
if not only or "packed" in only:
    for scheme, n, group, prg, nkeys in (("dpf", 32, "bytes", "aes128_mmo", 8500), ("halftree", 17, "u64", "chacha", 777)):
        p = Params(scheme=scheme, in_bits=n, group=group, prg=prg)
        s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=n + 1)
        o = orc.gen(p, s0s, alphas, betas, threads=4)
        oc, ooc = o if scheme == "halftree" else (o, None)
        ctx = fss_b200.Context(scheme, n, group, p.mod, prg, p.pred, p.prg_key, p.hash_key, p.in_bytes)
        os.environ["FSSB200_PACK_THREADS"] = "6"
        ctx.reserve_host(4096)
        rows = ctx.pack_rows(torch.from_numpy(oc.view(np.int32)))
        w = orc.eval(p, 1, s0s[:, 1], oc, xs, ooc, threads=4)
        ys = ctx.eval_packed(1, t(s0s[:, 1]), rows.to(dev), xs, None if ooc is None else t(ooc))
        same(f"eval_packed {scheme}-{n}", ys, w)
        yh = ctx.eval(1, torch.from_numpy(np.ascontiguousarray(s0s[:, 1]).view(np.int32)), torch.from_numpy(oc.view(np.int32)),
                      xs, None if ooc is None else torch.from_numpy(ooc.view(np.int32)))
        same(f"eval_host packed pipeline {scheme}-{n}", yh, w)
if not only or "walk" in only:
    # Grotto O(n) walk (TMA tiles, mode 4) and every mode of the host pipeline (pinned + pageable, ragged pieces)
    for n, nkeys, prg in ((32, 1000, "aes128_mmo"), (9, 77, "chacha"), (64, 333, "aes128_mmo")):
        p = Params(scheme="grotto", in_bits=n, prg=prg)
        s0s, alphas, _, xs = synth_inputs(p, nkeys, seed=n)
        ctx = fss_b200.Context("grotto", n, "bytes", prg=prg, prg_key=p.prg_key)
        cws = ctx.gen(t(s0s), alphas, None)
        w0 = ctx.grotto_walk(0, t(s0s[:, 0]), cws, xs).cpu().numpy()
        w1 = ctx.grotto_walk(1, t(s0s[:, 1]), cws, xs).cpu().numpy()
        assert np.array_equal(w0 ^ w1, np.array([int(a) <= int(x) for a, x in zip(alphas, xs)], np.uint8))
        print(f"ok grotto walk n={n} {prg}", flush=True)
if not only or "pipeline" in only:
    p = Params(scheme="dpf", in_bits=32)
    nk = 40000
    s0s, alphas, betas, xs = synth_inputs(p, nk, seed=31)
    oc = orc.gen(p, s0s, alphas, betas, threads=4)
    w = orc.eval(p, 0, s0s[:, 0], oc, xs, threads=4)
    ctx = fss_b200.Context("dpf", 32, "bytes", prg="aes128_mmo")
    ctx.reserve_host(9000)
    hs, hc = torch.from_numpy(np.ascontiguousarray(s0s[:, 0]).view(np.int32)), torch.from_numpy(oc.view(np.int32))
    hx = ctx.in_tensor(xs, torch.device("cpu"))
    for mode in (0, 1, 2, 3):
        ctx.set_host_mode(mode)
        for pin in (True, False):
            a = [v.pin_memory() for v in (hs, hc, hx)] if pin else [hs, hc, hx]
            same(f"host pipeline mode {mode} pinned={pin}", ctx.eval(0, a[0], a[1], a[2]), w)
torch.cuda.synchronize()
print("ALL OK")
