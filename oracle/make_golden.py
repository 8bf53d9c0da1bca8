#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/golden_v1.npz + golden_v1.json from the
UNMODIFIED reference (oracle/_ref/libfssref.so, built by `make -C oracle ref` from
/root/reference/include).  Run in the build container (the reference does not exist on the GPU box):

    make -C oracle ref && python oracle/make_golden.py

Every case stores its inputs (seeds, alphas, betas, xs) and the reference's outputs (cws, ocws,
ys of both parties, EvalAll outputs or their SHA-256 for large domains), so the tests can replay
it against the C restatement (CPU) and against the CUDA path (GPU) without the reference.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import HASH_KEY_BENCH, HASH_KEY_SAMPLE, Params, Ref, synth_inputs  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# samples/dpf_dcf_cpu.cu:39-55,98-116 fixture
FIX_SEEDS = np.array([[[0x11111111, 0x22222222, 0x33333333, 0x44444440],
                       [0x55555555, 0x66666666, 0x77777777, 0x88888880]]], dtype=np.uint32)
FIX_BETA = np.array([[7, 0, 0, 0]], dtype=np.uint32)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = Ref()
    arrays, manifest = {}, []

    def add_case(name, p: Params, s0s, alphas, betas, xs, evalall="full", evalall_keys=2, note=""):
        r = ref.gen(p, s0s, alphas, betas)
        cws, ocws = r if p.scheme == "halftree" else (r, None)
        case = {"name": name, "scheme": p.scheme, "in_bits": p.in_bits, "in_bytes": p.in_bytes, "group": p.group,
                "mod": str(p.mod), "prg": p.prg, "pred": p.pred, "prg_key": p.prg_key.hex(),
                "hash_key": p.hash_key.hex(), "alphas": [str(a) for a in alphas], "xs": [str(x) for x in xs],
                "note": note, "evalall": evalall}
        arrays[f"{name}/s0s"] = s0s
        arrays[f"{name}/betas"] = betas
        arrays[f"{name}/cws"] = cws
        if ocws is not None:
            arrays[f"{name}/ocws"] = ocws
        for party in (0, 1):
            if p.scheme != "grotto":
                arrays[f"{name}/ys{party}"] = ref.eval(p, party, s0s[:, party], cws, xs, ocws)
            if evalall != "none":
                k = min(evalall_keys, len(s0s))
                ya = ref.evalall(p, party, s0s[:k, party], cws[:k], None if ocws is None else ocws[:k])
                if evalall == "full":
                    arrays[f"{name}/all{party}"] = ya
                else:  # sha: per-key digest + a few sampled leaves
                    case[f"all{party}_sha256"] = [sha(ya[i]) for i in range(k)]
                    idx = np.array([0, 1, 2, (1 << p.in_bits) // 3, (1 << p.in_bits) - 1], dtype=np.int64)
                    case["all_sample_idx"] = idx.tolist()
                    arrays[f"{name}/all{party}_sample"] = ya[:, idx]
                case["evalall_keys"] = k
                if p.scheme == "grotto" and p.in_bits <= 12:
                    pt = ref.grotto_preprocess(p, party, s0s[:k, party], cws[:k])
                    arrays[f"{name}/pt{party}"] = pt
                    arrays[f"{name}/lookup{party}"] = ref.grotto_lookup(p, pt, xs[:k])
        manifest.append(case)

    # ---- PRG known answers --------------------------------------------------
    rng = np.random.default_rng(7)
    seeds = np.concatenate([FIX_SEEDS[0], rng.integers(0, 2 ** 32, size=(62, 4), dtype=np.uint64).astype(np.uint32)])
    arrays["prg/seeds"] = seeds
    for prg in ("aes128_mmo", "chacha"):
        for mul in (1, 2, 4):
            arrays[f"prg/{prg}_{mul}"] = ref.prg_gen(Params(prg=prg), mul, seeds)
    # Aes128Mmo == Aes128MmoRaw == Aes128Soft
    p_aes = Params(prg="aes128_mmo")
    assert np.array_equal(ref.prg_gen(p_aes, 2, seeds, prg_tag=2), arrays["prg/aes128_mmo_2"])
    assert np.array_equal(ref.prg_gen(p_aes, 2, seeds, prg_tag=3), arrays["prg/aes128_mmo_2"])

    # ---- C1: samples/dpf_dcf_cpu.cu verbatim (n=8, Bytes, AES-MMO, alpha=42, beta={7,0,0,0}) --
    for scheme in ("dpf", "dcf"):
        add_case(f"c1_{scheme}_n8_bytes_aes", Params(scheme=scheme, in_bits=8, in_bytes=1), FIX_SEEDS, [42], FIX_BETA,
                 [42], note="samples/dpf_dcf_cpu.cu fixture; ys at x=42")
        for x in (100, 10, 200):
            add_case(f"c1_{scheme}_n8_bytes_aes_x{x}", Params(scheme=scheme, in_bits=8, in_bytes=1), FIX_SEEDS, [42],
                     FIX_BETA, [x], evalall="none")
    add_case("c1_halftree_n8_bytes_aes", Params(scheme="halftree", in_bits=8, in_bytes=1, hash_key=HASH_KEY_SAMPLE),
             FIX_SEEDS, [42], FIX_BETA, [42], note="samples/half_tree_dpf_cpu.cu fixture")
    add_case("c1_halftree_n8_bytes_aes_x100",
             Params(scheme="halftree", in_bits=8, in_bytes=1, hash_key=HASH_KEY_SAMPLE), FIX_SEEDS, [42], FIX_BETA,
             [100], evalall="none")
    add_case("c1_grotto_n8_aes", Params(scheme="grotto", in_bits=8, in_bytes=1), FIX_SEEDS, [42], None, [42],
             note="SURVEY 8c grotto vector")
    # survey KATs at n=32 / n=64 with the same fixture
    add_case("kat_dpf_n32_bytes_aes", Params(scheme="dpf", in_bits=32), FIX_SEEDS, [42], FIX_BETA, [42],
             evalall="none")
    add_case("kat_dpf_n32_bytes_aes_xdeadbeef", Params(scheme="dpf", in_bits=32), FIX_SEEDS, [42], FIX_BETA,
             [0xDEADBEEF], evalall="none")
    add_case("kat_dcf_n64_u127_aes_x10", Params(scheme="dcf", in_bits=64, group="u128"), FIX_SEEDS, [42], FIX_BETA,
             [10], evalall="none")
    add_case("kat_dcf_n64_u127_aes_x2p40", Params(scheme="dcf", in_bits=64, group="u128"), FIX_SEEDS, [42], FIX_BETA,
             [1 << 40], evalall="none")

    # ---- BASELINE configs on seeded random keys ------------------------------------------------
    def rand_case(name, p, k, evalall="none", evalall_keys=2, seed=42):
        s0s, alphas, betas, xs = synth_inputs(p, k, seed=seed)
        n = p.in_bits
        if k >= 8:  # domain edges
            xs[1], xs[2], alphas[3], xs[3], alphas[4], xs[4] = 0, (1 << n) - 1, 0, 0, (1 << n) - 1, (1 << n) - 1
            xs[5] = max(alphas[5] - 1, 0)
            xs[6] = min(alphas[6] + 1, (1 << n) - 1)
        add_case(name, p, s0s, alphas, betas if p.scheme != "grotto" else None, xs, evalall, evalall_keys)

    for prg in ("aes128_mmo", "chacha"):
        t = "aes" if prg.startswith("aes") else "chacha"
        rand_case(f"c2_dpf_n32_bytes_{t}", Params(scheme="dpf", in_bits=32, prg=prg), 32)
        rand_case(f"c3_dcf_n64_u127_{t}", Params(scheme="dcf", in_bits=64, group="u128", prg=prg), 32)
        rand_case(f"c3_dcf_n32_u64_{t}_gt", Params(scheme="dcf", in_bits=32, group="u64", prg=prg, pred="gt"), 16)
        rand_case(f"c5_halftree_n32_bytes_{t}",
                  Params(scheme="halftree", in_bits=32, prg=prg, hash_key=HASH_KEY_BENCH), 32)
        rand_case(f"c4_dpf_n12_bytes_{t}", Params(scheme="dpf", in_bits=12, prg=prg), 8, "full", 2)
        rand_case(f"c4_dpf_n20_bytes_{t}", Params(scheme="dpf", in_bits=20, prg=prg), 8, "sha", 2)
        rand_case(f"c4_dcf_n12_u64_{t}", Params(scheme="dcf", in_bits=12, group="u64", prg=prg), 8, "full", 2)
        rand_case(f"c4_halftree_n12_bytes_{t}",
                  Params(scheme="halftree", in_bits=12, prg=prg, hash_key=HASH_KEY_BENCH), 8, "full", 2)
        rand_case(f"c4_halftree_n20_u64_{t}",
                  Params(scheme="halftree", in_bits=20, group="u64", prg=prg, hash_key=HASH_KEY_BENCH), 8, "sha", 1)
        rand_case(f"c5_grotto_n12_{t}", Params(scheme="grotto", in_bits=12, prg=prg), 8, "full", 2)
        rand_case(f"c5_grotto_n20_{t}", Params(scheme="grotto", in_bits=20, prg=prg), 8, "sha", 1)
    # the full-size C4 domain, one key, AES only (2^24: ~1 s per party on a core; 2^28: ~20 s)
    rand_case("c4_dpf_n24_bytes_aes", Params(scheme="dpf", in_bits=24), 8, "sha", 1)
    rand_case("c4_dpf_n28_bytes_aes", Params(scheme="dpf", in_bits=28), 8, "sha", 1)

    # ---- group / domain edge coverage (src/group_test.cu's 11 group types) -----------------------
    groups = [("bytes", 0), ("u8", 0), ("u8", 251), ("u16", 0), ("u16", 65521), ("u32", 0), ("u32", 4294967291),
              ("u64", 0), ("u64", 18446744073709551557), ("u128", 1 << 127), ("u128", (1 << 127) - 1),
              ("u128", (0x1234567812345678 << 64) | 0x9ABCDEF09ABCDEF1)]
    for g, mod in groups:
        for scheme in ("dpf", "dcf", "halftree"):
            for prg in ("aes128_mmo", "chacha"):
                t = "aes" if prg.startswith("aes") else "chacha"
                p = Params(scheme=scheme, in_bits=16, group=g, mod=mod, prg=prg, hash_key=HASH_KEY_BENCH)
                rand_case(f"grp_{scheme}_n16_{g}_{mod % 997}_{t}", p, 8, "none", seed=16 + len(g))
    for n in (1, 2, 3, 5, 33, 48, 100, 128):
        for scheme in ("dpf", "dcf", "halftree", "grotto"):
            p = Params(scheme=scheme, in_bits=n, group="bytes" if scheme == "grotto" else "u64",
                       hash_key=HASH_KEY_BENCH)
            rand_case(f"edge_{scheme}_n{n}", p, 8, "full" if n <= 5 else "none", 8, seed=n)

    os.makedirs(OUT_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(OUT_DIR, "golden_v1.npz"), **arrays)
    with open(os.path.join(OUT_DIR, "golden_v1.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "reference_commit": "c1ebc87 (v1.2.0)",
                   "cases": manifest}, f, indent=1)
    print(f"{len(manifest)} cases, {len(arrays)} arrays ->", OUT_DIR)


if __name__ == "__main__":
    main()
