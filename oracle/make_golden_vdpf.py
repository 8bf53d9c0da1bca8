#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/golden_vdpf_v1.{npz,json} from the UNMODIFIED
reference `fss::Vdpf` (oracle/_ref/libfssref.so = oracle/ref_vdpf.cpp over /root/reference/include):

    make -C oracle ref && python oracle/make_golden_vdpf.py

Each case stores inputs (seeds, alphas, betas, xs, hash IVs) and the reference's outputs: cws, cs, ocws,
Gen status, per-party (y, pi_tilde) of Eval, Prove over all points of a key, and EvalAll (ys + proof,
or SHA-256 of ys + proof for n = 16).  Also known answers of both hash interfaces.  A second file,
golden_vdpf_sha256_v1.*, holds the same for XorHash = Hash = fss::hash::Sha256 (hash/sha256.cuh).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import Params, Ref, synth_inputs  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
# samples/vdpf_cpu.cu / src/vdpf_test.cu:35-36 fixture seeds (same as the DPF sample)
FIX_SEEDS = np.array([[[0x11111111, 0x22222222, 0x33333333, 0x44444440],
                       [0x55555555, 0x66666666, 0x77777777, 0x88888880]]], dtype=np.uint32)
FIX_BETA = np.array([[7, 0, 0, 0]], dtype=np.uint32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build(hash_name):
    """All cases with XorHash = Hash = `hash_name` ("blake3": hash/blake3.cuh, "sha256": hash/sha256.cuh)."""
    ref = Ref()
    arrays, manifest = {}, []
    rng = np.random.default_rng(11)
    p0 = Params(scheme="vdpf", in_bits=8, hash=hash_name)
    sha_only = hash_name != "blake3"
    pre = "sha_" if sha_only else ""
    arrays["hash/xor_in"] = rng.integers(0, 2 ** 32, size=(32, 2, 4), dtype=np.uint64).astype(np.uint32)
    arrays["hash/hash_in"] = rng.integers(0, 2 ** 32, size=(32, 4, 4), dtype=np.uint64).astype(np.uint32)
    arrays["hash/xor_out"] = ref.hash(p0, 0, arrays["hash/xor_in"])
    arrays["hash/hash_out"] = ref.hash(p0, 1, arrays["hash/hash_in"])

    def add_case(name, p, s0s, alphas, betas, xs, evalall="none", evalall_keys=2):
        name = pre + name
        p.hash = hash_name
        assert ref.vdpf_supported(p), name
        cws, cs, ocws, status = ref.vdpf_gen(p, s0s, alphas, betas)
        assert not status.any()
        case = {"name": name, "in_bits": p.in_bits, "in_bytes": p.in_bytes, "group": p.group, "mod": str(p.mod),
                "prg": p.prg, "prg_key": p.prg_key.hex(), "hash_iv": bytes(p.hash_iv).hex(), "hash": hash_name,
                "alphas": [str(a) for a in alphas], "xs": [str(x) for x in xs], "evalall": evalall}
        arrays[f"{name}/s0s"], arrays[f"{name}/betas"] = s0s, betas
        arrays[f"{name}/cws"], arrays[f"{name}/cs"], arrays[f"{name}/ocws"] = cws, cs, ocws
        for party in (0, 1):
            ys, pis = ref.vdpf_eval(p, party, s0s[:, party], cws, cs, ocws, xs)
            arrays[f"{name}/ys{party}"], arrays[f"{name}/pis{party}"] = ys, pis
            # Prove: every key accumulates the hashes of ALL keys' points evaluated under its own cs is not
            # meaningful; use m = 4 consecutive points of the batch per key instead (shape [K/4, 4, 4, 4])
            k4 = len(s0s) // 4
            if k4:
                arrays[f"{name}/prove{party}"] = ref.vdpf_prove(p, pis[:4 * k4].reshape(k4, 4, 4, 4), cs[:k4])
            if evalall != "none":
                k = min(evalall_keys, len(s0s))
                ya, pa = ref.vdpf_evalall(p, party, s0s[:k, party], cws[:k], cs[:k], ocws[:k])
                arrays[f"{name}/allpi{party}"] = pa
                if evalall == "full":
                    arrays[f"{name}/all{party}"] = ya
                else:
                    case[f"all{party}_sha256"] = [sha(ya[i]) for i in range(k)]
                case["evalall_keys"] = k
        manifest.append(case)

    add_case("fix_vdpf_n8_bytes_aes", Params(scheme="vdpf", in_bits=8, in_bytes=1), FIX_SEEDS, [42], FIX_BETA, [42],
             "full", 1)
    add_case("fix_vdpf_n8_bytes_chacha", Params(scheme="vdpf", in_bits=8, in_bytes=1, prg="chacha"), FIX_SEEDS, [42],
             FIX_BETA, [100], "full", 1)

    def rand_case(name, p, k, evalall="none", evalall_keys=2, seed=42):
        s0s, alphas, betas, xs = synth_inputs(p, k, seed=seed)
        n = p.in_bits
        if k >= 8:
            xs[1], xs[2], alphas[3], xs[3], alphas[4], xs[4] = 0, (1 << n) - 1, 0, 0, (1 << n) - 1, (1 << n) - 1
        add_case(name, p, s0s, alphas, betas, xs, evalall, evalall_keys)

    for prg in ("aes128_mmo", "chacha"):
        t = "aes" if prg.startswith("aes") else "chacha"
        rand_case(f"vdpf_n32_bytes_{t}", Params(scheme="vdpf", in_bits=32, prg=prg), 32)
        rand_case(f"vdpf_n64_u127_{t}", Params(scheme="vdpf", in_bits=64, group="u128", prg=prg), 16)
        if not sha_only:  # (the reference shim instantiates fewer domains with Sha256)
            rand_case(f"vdpf_n20_u64_{t}", Params(scheme="vdpf", in_bits=20, group="u64", prg=prg), 16)
        rand_case(f"vdpf_n12_u32_{t}", Params(scheme="vdpf", in_bits=12, group="u32", prg=prg), 8, "full", 2)
        if not sha_only:
            rand_case(f"vdpf_n16_bytes_{t}", Params(scheme="vdpf", in_bits=16, prg=prg), 8, "sha", 2)
        rand_case(f"vdpf_n128_u64p_{t}", Params(scheme="vdpf", in_bits=128, group="u64", mod=18446744073709551557,
                                                 prg=prg), 8)
        for n in ((1,) if sha_only else (1, 3, 40)):
            rand_case(f"vdpf_n{n}_u64_{t}", Params(scheme="vdpf", in_bits=n, group="u64", prg=prg), 8,
                      "full" if n <= 3 else "none", 8, seed=n)

    return arrays, manifest


def main():
    for hash_name, stem in (("blake3", "golden_vdpf_v1"), ("sha256", "golden_vdpf_sha256_v1")):
        arrays, manifest = build(hash_name)
        if hash_name != "blake3":  # the hash known answers of this file live under their own keys
            arrays = {(k.replace("hash/", "hash_sha256/") if k.startswith("hash/") else k): v for k, v in arrays.items()}
        np.savez_compressed(os.path.join(OUT_DIR, stem + ".npz"), **arrays)
        with open(os.path.join(OUT_DIR, stem + ".json"), "w") as f:
            json.dump({"generator": "oracle/make_golden_vdpf.py", "reference_commit": "c1ebc87 (v1.2.0)",
                       "cases": manifest}, f, indent=1)
        print(f"{hash_name}: {len(manifest)} cases, {len(arrays)} arrays ->", OUT_DIR)


if __name__ == "__main__":
    main()
