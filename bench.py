#!/usr/bin/env python
"""bench.py -- throughput of the DPF/DCF PRG-tree hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --gpus N --single-process        # one process drives N GPUs (fssb200_*_multi), no torchrun

Workload (config.workload): BASELINE.json configs[1] -- batched DPF Eval, n = 32, one input per key,
group::Bytes, AES-128 MMO PRG, 2^22 independent keys PER GPU (weak scaling: keys are independent, every
rank evaluates its own key range, no data-path collective).  A step = one pass of the hot path over the
rank's batch = one launch of the point-evaluation kernel.  Inputs are synthetic (torch RNG on the device,
keys produced by this library's own Gen kernel, which the -m gpu tests pin to the reference's Gen) and
4.4 GB per GPU, i.e. far larger than the 126 MB L2.

Every timed leg is CHECKED before its number is reported: a sample of the timed output is compared bit for bit with
the oracle (checker role: oracle/ is never on the measured path), sharded legs are compared with the unsharded result
(SHA-256 of every rank's leaf range; NCCL gather of point outputs against a single-rank run).  Any mismatch exits
non-zero.

The one JSON line carries: `e2e` (the same metric through the host-buffer C-ABI entry point fssb200_eval_host with
pinned host buffers in the reference layout, copies inside the timed region, plus the box's concurrent H2D ceiling
measured in the same run), `roofline` (shared-memory lookup pipe, with the section-8d integer-ALU and HBM figures and
the ncu pipe utilisation of the committed capture beside it; `roofline.kernels` holds every other BASELINE config
and SURVEY section 8(f) row with its own fraction), `strong_scaling` (2^22 keys in total over the N GPUs),
`cpu_baseline` (the reference's own Eval with its OpenSSL AES-NI PRG on the host cores, rank 0, N=1 only), `clocks`.

`--impl reference` times the reference's CPU implementation of the same path (oracle/_ref, built from the
unmodified reference headers; falls back to the plain-C port) on all host threads, on the full 2^22-key batch.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DPF evals/sec (n=32, 2^22 keys, group::Bytes, AES-128 MMO)"
UNIT = "evals/s"
N_BITS = 32
KEYS_PER_GPU = 1 << 22
# algorithmic integer work per unit (fixed constants, independent of implementation)
OPS_PER_AES = 444
OPS_PER_EVAL_C2 = 32 * (OPS_PER_AES + 12)            # 14 592
OPS_PER_EVAL_C3 = 64 * (2 * OPS_PER_AES + 28)        # 58 624
OPS_PER_LEAF_C4 = 2 * OPS_PER_AES + 10               # 898
LOOKUPS_PER_AES = 160                                # shared-memory 32-bit table lookups per block (section 8d: L = 160)
BYTES_PER_EVAL_C2 = 1092
CHECK_KEYS = 1 << 12                                 # sample of every timed point-evaluation leg checked vs the oracle
CHECK_LEAVES = 1 << 16                               # ... of every full-domain leg


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (never touches the GPU or fss_b200)
# ---------------------------------------------------------------------------------------------------------

def cpu_engine():
    from oracle import Orc, Ref
    return Ref() if Ref.available() else Orc()


def synth_c2(nkeys: int, seed: int):
    """Seeded C2 inputs (numpy, no Python loops): clamped seeds / betas, uniform 32-bit alpha / x,
    every 16th x forced to alpha."""
    import numpy as np
    rng = np.random.default_rng(seed)
    s0s = rng.integers(0, 2 ** 32, size=(nkeys, 2, 4), dtype=np.uint64).astype(np.uint32)
    s0s[:, :, 3] &= 0xFFFFFFFE
    betas = rng.integers(0, 2 ** 32, size=(nkeys, 4), dtype=np.uint64).astype(np.uint32)
    betas[:, 3] &= 0xFFFFFFFE
    alphas = rng.integers(0, 2 ** 32, size=nkeys, dtype=np.uint64).astype(np.uint32)
    xs = rng.integers(0, 2 ** 32, size=nkeys, dtype=np.uint64).astype(np.uint32)
    xs[::16] = alphas[::16]
    return s0s, alphas, betas, xs


def cpu_dpf_eval_rate(engine, nkeys: int, threads: int, repeats: int = 1, prg: str = "aes128_mmo"):
    """Reference Dpf::Eval (dpf.cuh:170-214, Aes128Mmo<2> = OpenSSL AES-NI; prg="aes128_mmo_raw": Aes128MmoRaw<2>, the
    reference's own AES-NI intrinsics, SURVEY.md section 8d's "fair AES-NI ceiling") over nkeys keys on `threads` host
    threads, one PRG context set per thread.  Returns (evals/s, seconds per pass)."""
    import numpy as np
    from oracle import Params
    p = Params(scheme="dpf", in_bits=N_BITS, group="bytes", prg=prg)
    s0s, alphas, betas, xs = synth_c2(nkeys, 42)
    cws = engine.gen(p, s0s, alphas, betas, threads=threads)
    seeds = np.ascontiguousarray(s0s[:, 0])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        engine.eval(p, 0, seeds, cws, xs, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return nkeys / best, best


def cpu_other_rates(engine, threads: int) -> dict:
    """The reference's CPU path for the other BASELINE configs, timed beside the GPU rows of `roofline.kernels` (SURVEY.md
    section 8d "CPU baseline timing"): C3 = Dcf::Eval n=64 Uint<u128, 2^127> Aes128Mmo<4> over a bounded key sample, C4 =
    Dpf::EvalAll at the CPU-sized n=24 (2^28 leaves take ~16 s per key and core), one key per thread, one PRG context set per
    thread.  Numpy inputs only: no Python integer conversion inside the timed calls."""
    import numpy as np
    from oracle import Params
    out = {}
    rng = np.random.default_rng(7)

    def keys(p, k):
        s0s = rng.integers(0, 2 ** 32, size=(k, 2, 4), dtype=np.uint64).astype(np.uint32)
        s0s[:, :, 3] &= 0xFFFFFFFE
        betas = rng.integers(0, 2 ** 32, size=(k, 4), dtype=np.uint64).astype(np.uint32)
        betas[:, 3] &= 0xFFFFFFFE
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[p.in_bytes]
        hi = 1 << p.in_bits
        alphas = rng.integers(0, hi, size=k, dtype=np.uint64).astype(dt)
        xs = rng.integers(0, hi, size=k, dtype=np.uint64).astype(dt)
        xs[::16] = alphas[::16]
        return s0s, alphas, betas, xs

    def best_of(fn, n=2):
        best = None
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        return best

    p3 = Params(scheme="dcf", in_bits=64, group="u128", prg="aes128_mmo")
    k3 = 1 << 16
    s0s, alphas, betas, xs = keys(p3, k3)
    cws = engine.gen(p3, s0s, alphas, betas, threads=threads)
    seeds = np.ascontiguousarray(s0s[:, 0])
    t3 = best_of(lambda: engine.eval(p3, 0, seeds, cws, xs, threads=threads))
    out["c3_dcf_eval_n64_u127"] = {"value": k3 / t3, "unit": "evals/s", "cores": threads,
                                   "sample": f"{k3} keys, reference Dcf::Eval with Aes128Mmo<4>, best of 2 passes ({t3:.2f} s each)"}
    p4 = Params(scheme="dpf", in_bits=24, group="bytes", prg="aes128_mmo")
    k4 = max(1, min(threads, 4))
    s0s, alphas, betas, _ = keys(p4, k4)
    cws = engine.gen(p4, s0s, alphas, betas, threads=k4)
    seeds = np.ascontiguousarray(s0s[:, 0])
    t4 = best_of(lambda: engine.evalall(p4, 0, seeds, cws, threads=k4), 1)
    out["c4_dpf_evalall_n24"] = {"value": k4 * (1 << 24) / t4, "unit": "leaves/s", "cores": k4,
                                 "sample": f"{k4} keys x 2^24 leaves (the CPU-sized domain of SURVEY.md section 8d), reference "
                                           f"Dpf::EvalAll, one key per thread, one pass ({t4:.2f} s, includes allocating the output)"}
    return out


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import Params
    eng = cpu_engine()
    threads = host_threads()
    p = Params(scheme="dpf", in_bits=N_BITS, group="bytes", prg="aes128_mmo")
    # The arm runs the FULL 2^22-key batch of the own arm's config per step.  Only if that would take the whole run
    # past ~6 minutes (a host with very few cores) is the per-step sample cut to a power of two that fits.
    rate, _ = cpu_dpf_eval_rate(eng, 1 << 14, threads, 2)
    budget_s = 360.0 / max(1, args.steps + args.warmup)
    sample = KEYS_PER_GPU
    while sample > (1 << 14) and sample / rate > budget_s:
        sample >>= 1
    s0s, alphas, betas, xs = synth_c2(sample, 42)
    cws = eng.gen(p, s0s, alphas, betas, threads=threads)
    seeds = np.ascontiguousarray(s0s[:, 0])
    for _ in range(args.warmup):
        eng.eval(p, 0, seeds, cws, xs, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.eval(p, 0, seeds, cws, xs, threads=threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = sample / dt
    sample_desc = (f"{sample} of the 2^22 keys per step (seeded, reference Gen on the host), "
                   f"{threads} OpenMP threads, one EVP_CIPHER_CTX set per thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "batched DPF Eval, n=32, 1 input per key, group::Bytes, AES-128 MMO "
                               "(BASELINE configs[1])", "in_bits": N_BITS, "keys_per_gpu": sample,
                   "keys_total": sample, "party": 0, "cw_layout": "key-major Dpf::Cw (reference layout)",
                   "implementation": "reference CPU path: Dpf::Eval with Aes128Mmo<2> (OpenSSL AES-NI)",
                   "sample": sample_desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": eng.kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed regions run (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None and self._thread is None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------------
# checker: the oracle, on samples of the timed outputs
# ---------------------------------------------------------------------------------------------------------

class Checker:
    """oracle/ in its checker role.  Every method raises SystemExit on a mismatch."""

    def __init__(self):
        from oracle import Orc, Ref
        self.orc = Orc()                                   # the plain-C port: evaluates leaf RANGES of big domains
        self.eng = Ref() if Ref.available() else self.orc  # the compiled reference where it is present
        self.kind = self.eng.kind
        self.done = []

    @staticmethod
    def _np(t, dtype=None):
        import numpy as np
        a = t.detach().cpu().contiguous().numpy()
        return a.view(dtype or (np.uint32 if a.dtype == np.int32 else a.dtype))

    def _params(self, ctx):
        from oracle import Params
        return Params(scheme=ctx.scheme, in_bits=ctx.in_bits, group=ctx.group, mod=ctx.mod, prg=ctx.prg, pred=ctx.pred,
                      prg_key=bytes(ctx.prg_key), hash_key=bytes(ctx.hash_key), in_bytes=ctx.in_bytes)

    def point(self, name, ctx, party, seeds, cws, xs, ys, ocws=None, count=CHECK_KEYS):
        """First / last `count`/2 keys of a timed point-evaluation batch against the oracle's Eval."""
        import numpy as np
        n = seeds.shape[0]
        h = max(1, min(count // 2, n // 2))
        for sl in (slice(0, h), slice(n - h, n)):
            x = self._np(xs[sl])
            x = x.view(np.uint64 if ctx.in_bytes == 8 else np.uint32).reshape(-1)
            want = self.eng.eval(self._params(ctx), party, self._np(seeds[sl]), self._np(cws[sl]), x,
                                 None if ocws is None else self._np(ocws[sl]), threads=min(8, host_threads()))
            if not np.array_equal(self._np(ys[sl]), want):
                raise SystemExit(f"bench: {name}: timed output differs from the oracle ({self.kind})")
        self.done.append(f"{name}: {2 * h} keys of the timed batch == oracle Eval")

    def leaves(self, name, ctx, party, seed, cws, ys_key, leaf_begin, ocw=None, count=CHECK_LEAVES):
        """`count` leaves of one key of a timed full-domain output (ys_key = that key's leaves from leaf_begin on)."""
        import numpy as np
        cnt = min(count, ys_key.shape[0])
        p = self._params(ctx)
        want = self.orc.evalall(p, party, self._np(seed).reshape(1, 4), self._np(cws).reshape(1, ctx.ncw, 8),
                                None if ocw is None else self._np(ocw).reshape(1, 4), leaf_begin=leaf_begin,
                                leaf_count=cnt, threads=min(8, host_threads()))[0]
        got = self._np(ys_key[:cnt], want.dtype)
        if not np.array_equal(got.reshape(want.shape), want):
            raise SystemExit(f"bench: {name}: timed full-domain output differs from the oracle (port)")
        self.done.append(f"{name}: {cnt} leaves from {leaf_begin} of the timed output == oracle EvalAll (port)")

    def gen(self, name, ctx, s0s, alphas, betas, cws, ocws=None, count=256):
        import numpy as np
        a = self._np(alphas[:count]).view(np.uint64 if ctx.in_bytes == 8 else np.uint32).reshape(-1)
        o = self.eng.gen(self._params(ctx), self._np(s0s[:count]), a, None if betas is None else self._np(betas[:count]),
                         threads=min(8, host_threads()))
        oc, ooc = o if ctx.scheme == "halftree" else (o, None)
        got = self._np(cws[:count]).copy()
        if ctx.scheme in ("dpf", "halftree", "grotto"):   # bytes 17..31 of {int4 s; bool} are padding (dpf.cuh:76-81)
            got.view(np.uint8).reshape(count, ctx.ncw, 32)[:, :, 17:] = 0
            oc = oc.copy()
            oc.view(np.uint8).reshape(count, ctx.ncw, 32)[:, :, 17:] = 0
            if ctx.scheme != "halftree":                  # entry n: only .s is defined
                got.view(np.uint8).reshape(count, ctx.ncw, 32)[:, -1, 16:] = 0
                oc.view(np.uint8).reshape(count, ctx.ncw, 32)[:, -1, 16:] = 0
        if not np.array_equal(got, oc) or (ooc is not None and not np.array_equal(self._np(ocws[:count]), ooc)):
            raise SystemExit(f"bench: {name}: generated keys differ from the oracle ({self.kind})")
        self.done.append(f"{name}: {count} keys == oracle Gen")


def sha256_tensor(t) -> bytes:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).digest()


# ---------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------

def run_own_arm(args) -> None:
    import torch
    import torch.distributed as dist

    import fss_b200  # fails loudly if libfssb200.so is missing
    from fss_b200.sharding import gather_point_outputs, key_shard, leaf_shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set; stdout must carry the ONE JSON line only,
        # so file descriptor 1 points at stderr while the communicator is created (first collective included)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(v: float, op="max") -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    def rand_i32(shape, gen):
        return torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int64, device=dev, generator=gen).to(torch.int32)

    def make_keys(ctx, nkeys, gen, wide=False, bits=None):
        s0s = rand_i32((nkeys, 2, 4), gen)
        betas = rand_i32((nkeys, 4), gen)
        s0s[:, :, 3] &= ~1
        betas[:, 3] &= ~1
        if wide:
            alphas = torch.randint(-2 ** 63, 2 ** 63 - 1, (nkeys,), dtype=torch.int64, device=dev, generator=gen)
            xs = torch.randint(-2 ** 63, 2 ** 63 - 1, (nkeys,), dtype=torch.int64, device=dev, generator=gen)
        else:
            alphas, xs = rand_i32((nkeys,), gen), rand_i32((nkeys,), gen)
        if bits is not None and bits < 32:
            alphas, xs = alphas & ((1 << bits) - 1), xs & ((1 << bits) - 1)
        xs[::16] = alphas[::16]  # exercise the beta branch
        r = ctx.gen(s0s, alphas, None if ctx.scheme == "grotto" else betas)
        return s0s, alphas, betas, xs, r

    def timed(fn, steps, warmup):
        """W warm-up steps, then K steps bracketed by barrier + synchronize; per-step CUDA events on the
        launching (torch current) stream.  Returns (ms per step max over ranks, mean kernel ms)."""
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_all0.record()
        for a, b in ev:
            a.record()
            fn()
            b.record()
        t_all1.record()
        barrier()
        total_ms = t_all0.elapsed_time(t_all1)
        per_launch = sum(a.elapsed_time(b) for a, b in ev) / steps
        return reduce_ranks(total_ms / steps), per_launch

    def wall(fn, steps, warmup=2):
        """Blocking host calls: wall clock around K calls, max over ranks (seconds per step)."""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return reduce_ranks((time.perf_counter() - t0) / steps)

    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    sampler = ClockSampler(local)
    chk = Checker() if not args.no_check else None
    nkeys = args.keys

    # ---- main metric: C2 ----------------------------------------------------------------------------------
    ctx = fss_b200.Context("dpf", N_BITS, "bytes", prg="aes128_mmo")
    s0s, alphas, betas, xs, cws = make_keys(ctx, nkeys, gen)
    seeds0 = s0s[:, 0].contiguous()
    ys = torch.empty((nkeys, 4), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    sampler.start()
    ms_step, ms_kernel = timed(lambda: ctx.eval(0, seeds0, cws, xs, out=ys), args.steps, args.warmup)
    sampler.stop()
    launches = ctx.launch_count() - l0 - args.warmup
    value = world * nkeys / (ms_step * 1e-3)
    # correctness guards on the timed batch: the oracle on a sample, reconstruction on all of it
    if chk:
        chk.gen("C2 keys (library Gen)", ctx, s0s, alphas, betas, cws)
        chk.point("C2 dpf n=32 bytes aes", ctx, 0, seeds0, cws, xs, ys)
    y1 = ctx.eval(1, s0s[:, 1].contiguous(), cws, xs)
    hit = (xs == alphas).unsqueeze(1)
    if not torch.equal(ys ^ y1, torch.where(hit, betas, torch.zeros_like(betas))):
        raise SystemExit("bench: DPF reconstruction check failed on the timed batch")
    del y1

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    peaks = {}
    if rank == 0:
        for kind, name in ((0, "lop3"), (1, "imad"), (2, "lop3_imad_mixed"), (3, "lds32_conflict_free"), (4, "prmt"),
                           (5, "idp4a"), (6, "prmt_idp4a_mixed")):
            peaks[name] = fss_b200.microbench(kind, local)
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(measured.get("hbm_gbs", 6650.0))
    traffic, ncu_pipes = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = prof.get("dpf_point_c2")
        ncu_pipes = prof.get("dpf_point_c2_ncu")
    except Exception:
        pass
    int_achieved = nkeys * OPS_PER_EVAL_C2 / (ms_kernel * 1e-3) / 1e12
    int_peak = (peaks.get("lop3", 0.0) / 1e12) or None
    lds_achieved = nkeys * N_BITS * LOOKUPS_PER_AES / (ms_kernel * 1e-3) / 1e12
    lds_peak = (peaks.get("lds32_conflict_free", 0.0) / 1e12) or None
    hbm_achieved = nkeys * BYTES_PER_EVAL_C2 / (ms_kernel * 1e-3) / 1e9
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    # The binding roofline of a T-table AES kernel is the shared-memory lookup pipe (32 conflict-free LDS.32 lanes per
    # clock and SM, measured on this box in this run and cross-checked against the architectural 32 x SMs x SM clock);
    # the integer-ALU roofline of SURVEY.md section 8d (canonical 444 ops per block against the LOP3 issue rate) and the
    # HBM figures are reported beside it.  Not HBM-bound: see `hbm`.
    roofline = {
        "bound": "smem_lsu", "achieved": lds_achieved, "peak": lds_peak, "unit": "Tlookups/s",
        "frac": (lds_achieved / lds_peak) if lds_peak else None, "traffic": traffic,
        "kernel": "point_kernel<DPF,Bytes,AES>", "ms_per_launch": ms_kernel,
        "algorithmic_lookups_per_eval": N_BITS * LOOKUPS_PER_AES,
        "peak_source": "on-box conflict-free ld.shared.u32 rate measured in this run (fssb200_microbench kind 3); "
                       "MEASURED_PEAKS.json has no shared-memory or integer peak",
        "peak_architectural": {"value": 32 * sm_count * 1965e6 / 1e12, "unit": "Tlookups/s",
                               "how": f"32 LDS.32 lanes/clk/SM x {sm_count} SMs x 1.965 GHz (sm_max_mhz of MEASURED_PEAKS.json)"},
        "int_alu": {"bound": "int_alu", "achieved": int_achieved, "peak": int_peak, "unit": "Tops/s(int32)",
                    "frac": (int_achieved / int_peak) if int_peak else None,
                    "algorithmic_ops_per_eval": OPS_PER_EVAL_C2,
                    "note": "canonical T-table count of SURVEY.md section 8d (444 ops per AES block + 12 glue per level) "
                            "against the measured LOP3 issue rate; this implementation issues ~240 ALU-pipe + ~90 "
                            "FMA-pipe integer instructions per level, so the fraction exceeds 1 and is not a bound -- "
                            "the honest issue-slot figures are in `ncu`"},
        "ncu": ncu_pipes,
        "microbench_ops_per_s": peaks,
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_achieved / hbm_peak, "algorithmic_bytes_per_eval": BYTES_PER_EVAL_C2,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if measured else "fallback 6.65 TB/s"},
        "aes_blocks_per_s": nkeys * 32 / (ms_kernel * 1e-3),
    }

    def lsu_frac(units, blocks_per_unit, ms):
        return (units * blocks_per_unit * LOOKUPS_PER_AES / (ms * 1e-3) / 1e12 / lds_peak) if lds_peak else None

    if world > 1:  # every rank needs the LDS peak for its own fractions
        t = torch.tensor([lds_peak or 0.0, int_peak or 0.0], dtype=torch.float64, device=dev)
        dist.broadcast(t, 0)
        lds_peak, int_peak = float(t[0]) or None, float(t[1]) or None

    # ---- e2e: host buffers through the C ABI -------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_seeds = seeds0.cpu().pin_memory()
        h_cws = cws.cpu().pin_memory()
        h_xs = xs.cpu().pin_memory()
        h_ys = torch.empty((nkeys, 4), dtype=torch.int32).pin_memory()
        ys_host = ys.cpu()
        e_steps = max(1, min(args.steps, 5))
        row_b = ctx.packed_row_bytes(local)
        in_b = h_seeds.shape[1] * 4 + h_xs.element_size()

        def e2e_leg(mode):
            ctx.set_host_mode(mode)
            dt = wall(lambda: ctx.eval(0, h_seeds, h_cws, h_xs, out=h_ys), e_steps)
            if not torch.equal(h_ys, ys_host):
                raise SystemExit(f"bench: host-buffer path (mode {mode}) disagrees with the device path")
            st = ctx.host_stats()
            h2d = st["packed_keys"] * (row_b + in_b) + st["direct_keys"] * (ctx.ncw * 32 + in_b)
            return {"value": world * nkeys / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h_ys.numel() * 4,
                    "packed_keys": st["packed_keys"], "direct_keys": st["direct_keys"], "host_threads": st["threads"]}

        # concurrent pinned-H2D ceiling of the box with all `world` ranks copying at once (what tools/h2d_probe.py
        # measures stand-alone): 1 GiB of the pinned key buffer, 6 back-to-back copies per rank
        probe_bytes = min(1 << 30, h_cws.numel() * 4)
        d_probe = torch.empty(probe_bytes, dtype=torch.uint8, device=dev)
        src = h_cws.view(-1).view(torch.uint8)[:probe_bytes]
        d_probe.copy_(src, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(6):
            d_probe.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        my_gbs = 6 * probe_bytes / (time.perf_counter() - t0) / 1e9
        ceiling = {"per_rank_min_gbs": reduce_ranks(my_gbs, "min"), "per_rank_max_gbs": reduce_ranks(my_gbs, "max"),
                   "total_gbs": reduce_ranks(my_gbs, "sum"), "ranks": world,
                   # every rank holds the same number of keys, so a step lasts as long as the SLOWEST link needs
                   "balanced_total_gbs": world * reduce_ranks(my_gbs, "min"),
                   "how": "every rank copies 1 GiB of pinned host memory to its GPU 6 times back to back, all ranks at "
                          "once (cudaMemcpyAsync, wall clock per rank)"}
        del d_probe
        e2e = e2e_leg(0)
        e2e["api"] = ("fssb200_eval_host: reference-layout keys (32-byte Dpf::Cw) in pinned host buffers, results back in "
                      "pinned host memory, wall clock around the blocking call; automatic mode: with 1-2 ranks per host the "
                      "adaptive pipeline (host threads strip the 15 padding bytes of each Cw for chunks taken from the front of "
                      "the batch while rows from the back cross the link as they are whenever it would otherwise idle), with "
                      ">= 3 ranks the rows cross as they are (the host memory system, not the links, is the bound)")
        e2e["direct_copy"] = e2e_leg(1)     # reference layout crosses the link as it is
        e2e["staged_only"] = e2e_leg(2)     # every chunk packed by the host threads
        if world >= 3:                      # automatic mode = direct from three ranks on: show what it is chosen over
            e2e["adaptive_forced"] = e2e_leg(3)
        ctx.set_host_mode(0)
        # link utilisation and the ceiling a plain copy of the reference layout could reach on this box
        agg_h2d_gbs = world * e2e["h2d_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9
        plain_ceiling = ceiling["total_gbs"] * 1e9 / (ctx.ncw * 32 + in_b)
        e2e["h2d_ceiling"] = ceiling
        e2e["h2d_gbs_total"] = agg_h2d_gbs
        e2e["link_utilisation"] = agg_h2d_gbs / ceiling["total_gbs"]
        e2e["frac_of_balanced_h2d_ceiling"] = agg_h2d_gbs / ceiling["balanced_total_gbs"]
        e2e["evals_per_s_if_reference_layout_at_ceiling"] = plain_ceiling
        e2e["value_over_plain_copy_ceiling"] = e2e["value"] / plain_ceiling
        e2e["direct_copy"]["frac_of_h2d_ceiling"] = (world * e2e["direct_copy"]["h2d_bytes_per_step"] /
                                                     (e2e["direct_copy"]["ms_per_step"] * 1e-3) / 1e9 / ceiling["total_gbs"])
        # the same keys in the compact level-major layout (fssb200_relayout: 16 B + 1 bit per level instead of the
        # 32-byte Dpf::Cw) through fssb200_eval_levelmajor_host
        lay = ctx.relayout(cws)
        torch.cuda.synchronize()
        ms_lm, msk_lm = timed(lambda: ctx.eval_levelmajor(0, seeds0, lay, xs, out=ys), max(3, min(args.steps, 10)), 3)
        h_lay = tuple(None if t is None else t.cpu().pin_memory() for t in lay)
        dt_lm = wall(lambda: ctx.eval_levelmajor(0, h_seeds, h_lay, h_xs, out=h_ys), e_steps)
        if not torch.equal(h_ys, ys_host):
            raise SystemExit("bench: level-major host path disagrees with the device path")
        e2e["compact_levelmajor"] = {
            "value": world * nkeys / dt_lm, "unit": UNIT, "ms_per_step": dt_lm * 1e3,
            "h2d_bytes_per_step": h_seeds.numel() * 4 + h_xs.numel() * 4 + sum(t.numel() * 4 for t in h_lay if t is not None),
            "d2h_bytes_per_step": h_ys.numel() * 4,
            "kernel_only_evals_per_s": world * nkeys / (ms_lm * 1e-3), "kernel_lsu_roofline_frac": lsu_frac(nkeys, 32, msk_lm),
            "api": "fssb200_eval_levelmajor_host: keys held by the caller in the level-major layout of fssb200_relayout "
                   "(not the reference's Cw layout; reported beside the drop-in number, not instead of it)"}
        del h_seeds, h_cws, h_xs, h_ys, h_lay, lay, ys_host

    # ---- strong scaling: BASELINE's "2^22 keys @ 1/2/4/8 B200" -- the SAME 2^22 keys split over the N GPUs ---------------
    strong = None
    if world > 1 and nkeys == KEYS_PER_GPU:
        b, e = key_shard(KEYS_PER_GPU, rank, world)
        ks = e - b
        ms_s, _ = timed(lambda: ctx.eval(0, seeds0[:ks], cws[:ks], xs[:ks], out=ys[:ks]), args.steps, args.warmup)
        strong = {"value": KEYS_PER_GPU / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s, "keys_total": KEYS_PER_GPU,
                  "keys_per_gpu": ks, "scaling": "strong",
                  "resident_keys_per_gpu_wave": sm_count * 768,
                  "note": "one kernel launch per GPU; waves of 148 SMs x 768 resident keys"}
    elif world == 1:
        strong = {"value": value, "unit": UNIT, "ms_per_step": ms_step, "keys_total": nkeys, "keys_per_gpu": nkeys,
                  "scaling": "strong", "note": "N = 1: identical to the main line"}

    # ---- multi-GPU parity: NCCL gather of sharded point outputs == single-rank evaluation ------------------------------
    if world > 1 and chk:
        gk = 1 << 16
        cg = torch.Generator(device=dev).manual_seed(999)          # same keys on every rank
        s_c, a_c, b_c, x_c, cw_c = make_keys(ctx, gk, cg)
        b, e = key_shard(gk, rank, world)
        y_mine = ctx.eval(0, s_c[b:e, 0].contiguous(), cw_c[b:e].contiguous(), x_c[b:e].contiguous())
        y_all = gather_point_outputs(y_mine, gk)
        y_one = ctx.eval(0, s_c[:, 0].contiguous(), cw_c, x_c)
        if not torch.equal(y_all, y_one):
            raise SystemExit("bench: NCCL gather of the key-sharded outputs differs from the single-rank evaluation")
        chk.point("sharded dpf n=32 (NCCL all_gather of every rank's key range)", ctx, 0, s_c[:, 0].contiguous(), cw_c, x_c,
                  y_all)
        del s_c, b_c, cw_c, y_all, y_one
    del cws, s0s, betas, ys
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs and the SURVEY section 8(f) rows ----------------------------------------------------------
    extra = {}
    if not args.no_extra:
        x_steps, x_warm = max(3, min(args.steps, 10)), 3
        sampler.start()
        # C3: DCF n=64, Uint<u128,2^127>, Aes128Mmo<4>
        k3 = args.keys
        c3 = fss_b200.Context("dcf", 64, "u128", prg="aes128_mmo")
        s0s, alphas, betas, xs, cws = make_keys(c3, k3, gen, wide=True)
        seeds0 = s0s[:, 0].contiguous()
        ys = torch.empty((k3, 4), dtype=torch.int32, device=dev)
        ms3, msk3 = timed(lambda: c3.eval(0, seeds0, cws, xs, out=ys), x_steps, x_warm)
        if chk:
            chk.point("C3 dcf n=64 u127 aes", c3, 0, seeds0, cws, xs, ys)
        extra["dcf_n64_u127_aes"] = {
            "value": world * k3 / (ms3 * 1e-3), "unit": "evals/s", "ms_per_step": ms3, "keys_per_gpu": k3,
            "int_roofline_frac": (k3 * OPS_PER_EVAL_C3 / (msk3 * 1e-3) / 1e12 / int_peak) if int_peak else None,
            "lsu_roofline_frac": lsu_frac(k3, 128, msk3)}
        kg = min(k3, 1 << 21)
        al3 = c3.in_tensor(alphas[:kg], dev)
        msg, mskg = timed(lambda: c3.gen(s0s[:kg], al3, betas[:kg]), x_steps, x_warm)
        if chk:
            chk.gen("gen dcf n=64 u127", c3, s0s, alphas, betas, cws)
        extra["gen_dcf_n64_u127"] = {"value": world * kg / (msg * 1e-3), "unit": "keys/s", "ms_per_step": msg, "keys_per_gpu": kg,
                                     "aes_blocks_per_key": 8 * 64, "lsu_roofline_frac": lsu_frac(kg, 8 * 64, mskg)}
        del s0s, betas, cws, ys, seeds0, al3
        torch.cuda.empty_cache()
        # C5: Half-Tree DPF n=32, 2^20 keys
        k5 = min(args.keys, 1 << 20)
        c5 = fss_b200.Context("halftree", 32, "bytes", prg="aes128_mmo")
        s0s, alphas, betas, xs, (cws, ocws) = make_keys(c5, k5, gen)
        seeds0 = s0s[:, 0].contiguous()
        ys = torch.empty((k5, 4), dtype=torch.int32, device=dev)
        ms5, msk5 = timed(lambda: c5.eval(0, seeds0, cws, xs, ocws, out=ys), x_steps, x_warm)
        if chk:
            chk.point("C5 halftree n=32 bytes aes", c5, 0, seeds0, cws, xs, ys, ocws)
            chk.gen("gen halftree n=32", c5, s0s, alphas, betas, cws, ocws)
        extra["halftree_n32_aes"] = {
            "value": world * k5 / (ms5 * 1e-3), "unit": "evals/s", "ms_per_step": ms5, "keys_per_gpu": k5,
            "lsu_roofline_frac": lsu_frac(k5, 32, msk5)}
        al5 = c5.in_tensor(alphas, dev)
        msg, mskg = timed(lambda: c5.gen(s0s, al5, betas), x_steps, x_warm)
        extra["gen_halftree_n32"] = {"value": world * k5 / (msg * 1e-3), "unit": "keys/s", "ms_per_step": msg, "keys_per_gpu": k5,
                                     "aes_blocks_per_key": 2 * 32 + 2, "lsu_roofline_frac": lsu_frac(k5, 2 * 32 + 2, mskg)}
        del s0s, betas, cws, ocws, ys, seeds0, al5
        torch.cuda.empty_cache()
        # C5, Grotto half (both alternatives):
        #  (i) O(n) walk, n = 32, 2^20 keys: reconstruction-equal, share-parity unpinned (include/fssb200.h)
        cgw = fss_b200.Context("grotto", 32, prg="aes128_mmo")
        s0s, alphas, _, xs, cws = make_keys(cgw, k5, gen)
        w0 = torch.empty((k5,), dtype=torch.uint8, device=dev)
        seeds0, seeds1 = s0s[:, 0].contiguous(), s0s[:, 1].contiguous()
        msw, mskw = timed(lambda: cgw.grotto_walk(0, seeds0, cws, xs, out=w0), x_steps, x_warm)
        w1 = cgw.grotto_walk(1, seeds1, cws, xs)
        u = lambda t: t.to(torch.int64) & 0xFFFFFFFF  # noqa: E731
        if not torch.equal((w0 ^ w1).bool(), u(alphas) <= u(xs)):
            raise SystemExit("bench: Grotto walk does not reconstruct to 1[alpha <= x] on the timed batch")
        if chk:
            chk.done.append(f"C5 grotto walk n=32: {k5} keys, share0 ^ share1 == 1[alpha <= x] on the whole timed batch")
        extra["grotto_walk_n32_aes"] = {
            "value": world * k5 / (msw * 1e-3), "unit": "evals/s", "ms_per_step": msw, "keys_per_gpu": k5,
            "aes_blocks_per_eval": 2 * 32 - 1, "lsu_roofline_frac": lsu_frac(k5, 2 * 32 - 1, mskw),
            "parity": "reconstruction-equal to GrottoDcf::Eval; per-share bits unpinned (SURVEY.md H6)"}
        del s0s, cws, w0, w1, seeds0, seeds1
        torch.cuda.empty_cache()
        #  (ii) batched Preprocess + Eval at the reference's benchmarked n = 20 (2 MiB parity tree per key), per-share bit-exact
        kp = 1 << 10
        cgp = fss_b200.Context("grotto", 20, prg="aes128_mmo")
        s0s, alphas, _, xs, cws = make_keys(cgp, kp, gen, bits=20)
        seeds0 = s0s[:, 0].contiguous()
        pt_box = {}

        def prep_eval():
            pt_box["pt"] = cgp.grotto_preprocess(0, seeds0, cws)
            pt_box["y"] = cgp.grotto_lookup(pt_box["pt"], xs)
        msp, mskp = timed(prep_eval, max(3, x_steps // 2), 2)
        if chk:
            import numpy as np
            from oracle import Params
            pg = Params(scheme="grotto", in_bits=20, prg="aes128_mmo")
            hs, hc = Checker._np(seeds0[:2]), Checker._np(cws[:2])
            want_pt = chk.eng.grotto_preprocess(pg, 0, hs, hc, threads=min(8, host_threads()))
            if not np.array_equal(Checker._np(pt_box["pt"][:2]), want_pt):
                raise SystemExit("bench: Grotto Preprocess differs from the oracle")
            hx = Checker._np(xs[:2]).view(np.uint32).reshape(-1)
            if not np.array_equal(Checker._np(pt_box["y"][:2]), chk.eng.grotto_lookup(pg, want_pt, hx)):
                raise SystemExit("bench: Grotto Eval differs from the oracle")
            chk.done.append("C5 grotto Preprocess+Eval n=20: parity trees and shares of 2 keys == oracle")
        extra["grotto_preprocess_eval_n20_aes"] = {
            "value": world * kp / (msp * 1e-3), "unit": "keys/s (Preprocess + 1 Eval each)", "ms_per_step": msp,
            "keys_per_gpu": kp, "leaves_per_s": world * kp * (1 << 20) / (msp * 1e-3),
            "lsu_roofline_frac": lsu_frac(kp * (1 << 20), 2.0, mskp), "parity": "bit-exact per share vs grotto_dcf.cuh:94-135",
            "tree_gib_per_gpu": kp * ((2 << 20) - 1) / 2 ** 30}
        pt_box.clear()
        del s0s, cws, seeds0
        torch.cuda.empty_cache()
        # C4: DPF EvalAll n=28, 64 keys over 8 GPUs = 8 keys (32 GiB of leaves) per GPU and step
        n4, k4 = args.evalall_bits, args.evalall_keys
        c4 = fss_b200.Context("dpf", n4, "bytes", prg="aes128_mmo")
        s0s, alphas, betas, xs, cws = make_keys(c4, k4, gen, bits=n4)
        seeds0 = s0s[:, 0].contiguous()
        out = torch.empty((k4, 1 << n4, 4), dtype=torch.int32, device=dev)
        ms4, msk4 = timed(lambda: c4.eval_all(0, seeds0, cws, out=out), max(2, x_steps // 2), 2)
        leaves = k4 * (1 << n4)
        if chk:
            chk.leaves("C4 dpf evalall", c4, 0, seeds0[k4 - 1], cws[k4 - 1], out[k4 - 1], 0)
            lb = (1 << n4) - CHECK_LEAVES
            chk.leaves("C4 dpf evalall (last leaves)", c4, 0, seeds0[0], cws[0], out[0, lb:], lb)
        extra["dpf_evalall"] = {
            "value": world * leaves / (ms4 * 1e-3), "unit": "leaves/s", "ms_per_step": ms4, "in_bits": n4,
            "keys_per_gpu": k4, "output_gib_per_gpu": leaves * 16 / 2 ** 30,
            "int_roofline_frac": (leaves * OPS_PER_LEAF_C4 / (msk4 * 1e-3) / 1e12 / int_peak) if int_peak else None,
            "lsu_roofline_frac": lsu_frac(leaves, 2.0, msk4),
            "hbm_write_gbs": leaves * 16 / (msk4 * 1e-3) / 1e9}
        # C4 as BASELINE configs[3] words it: "subtrees sharded" -- every rank expands ITS leaf range of every key
        # (8 keys per GPU in the job, so the per-GPU work stays 32 GiB of leaves: weak scaling), no collective
        if world > 1:
            kk = k4 * world
            s0s, alphas, betas, xs, cws = make_keys(c4, kk, torch.Generator(device=dev).manual_seed(4242), bits=n4)  # same keys on every rank
            seeds0 = s0s[:, 0].contiguous()
            lb, lc = leaf_shard(n4, c4.granule(), rank, world)
            outv = out.view(-1)[: kk * lc * 4].view(kk, lc, 4)
            ms4s, _ = timed(lambda: c4.eval_all(0, seeds0, cws, leaf_begin=lb, leaf_count=lc, out=outv), max(2, x_steps // 2), 2)
            if chk:
                # every rank: SHA-256 of ITS leaf range of key 0; rank 0: the same ranges of the unsharded EvalAll of key 0
                chk.leaves(f"C4 sharded (rank {rank})", c4, 0, seeds0[kk - 1], cws[kk - 1], outv[kk - 1], lb)
                mine = torch.frombuffer(bytearray(sha256_tensor(outv[0])), dtype=torch.uint8).to(dev)
                digests = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(digests, mine)
                if rank == 0:
                    del outv
                    full = out.view(-1)[: (1 << n4) * 4].view(1, 1 << n4, 4)
                    c4.eval_all(0, seeds0[:1], cws[:1], out=full)
                    for r in range(world):
                        rb, rc_ = leaf_shard(n4, c4.granule(), r, world)
                        if sha256_tensor(full[0, rb:rb + rc_]) != bytes(digests[r].cpu().tolist()):
                            raise SystemExit(f"bench: leaf range of rank {r} differs from the unsharded EvalAll")
                    chk.done.append(f"C4 sharded: SHA-256 of all {world} ranks' leaf ranges of key 0 == unsharded EvalAll")
            extra["dpf_evalall_subtree_sharded"] = {
                "value": kk * (1 << n4) / (ms4s * 1e-3), "unit": "leaves/s", "ms_per_step": ms4s, "in_bits": n4,
                "keys_total": kk, "leaf_range_per_gpu": [int(lb), int(lc)], "output_gib_per_gpu": kk * lc * 16 / 2 ** 30}
        del out, cws
        torch.cuda.empty_cache()
        if not args.no_frows:
            # ---- the callers / data formats either side of the path --------------------------
            # f-3: Half-Tree EvalAll n=28, DCF EvalAll n=24
            for name, scheme, n, group, blocks, k in (("evalall_halftree_n28_bytes", "halftree", 28, "bytes", 1.5, 4),
                                                      ("evalall_dcf_n24_u127", "dcf", 24, "u128", 4.0, 16),
                                                      ("evalall_dcf_n24_bytes", "dcf", 24, "bytes", 4.0, 16)):
                cf = fss_b200.Context(scheme, n, group, prg="aes128_mmo")
                s0s, alphas, betas, xs, r = make_keys(cf, k, gen, bits=n)
                cws, ocws = r if scheme == "halftree" else (r, None)
                seeds0 = s0s[:, 0].contiguous()
                out = torch.empty((k, 1 << n, 4), dtype=torch.int32, device=dev)
                msf, mskf = timed(lambda: cf.eval_all(0, seeds0, cws, ocws, out=out), max(2, x_steps // 2), 2)
                if chk:
                    chk.leaves(name, cf, 0, seeds0[k - 1], cws[k - 1], out[k - 1], 0, None if ocws is None else ocws[k - 1])
                lv = k * (1 << n)
                extra[name] = {"value": world * lv / (msf * 1e-3), "unit": "leaves/s", "ms_per_step": msf, "in_bits": n,
                               "keys_per_gpu": k, "aes_blocks_per_leaf": blocks, "lsu_roofline_frac": lsu_frac(lv, blocks, mskf)}
                del out, cws, ocws, s0s
                torch.cuda.empty_cache()
            # f-3: Grotto expand / EvalAll / Preprocess, n=26, 16 keys
            n, k = 26, 16
            cf = fss_b200.Context("grotto", n, prg="aes128_mmo")
            s0s, alphas, _, xs, cws = make_keys(cf, k, gen, bits=n)
            seeds0 = s0s[:, 0].contiguous()
            lv = k * (1 << n)
            box = {}
            for name, fn in (("grotto_expand_n26", lambda: box.__setitem__("t", cf.grotto_expand(0, seeds0, cws))),
                             ("grotto_evalall_n26", lambda: box.__setitem__("t", cf.eval_all(0, seeds0, cws))),
                             ("grotto_preprocess_n26", lambda: box.__setitem__("t", cf.grotto_preprocess(0, seeds0, cws)))):
                box.clear()
                torch.cuda.empty_cache()
                msf, mskf = timed(fn, 3, 2)
                if chk and name == "grotto_evalall_n26":
                    chk.leaves(name, cf, 0, seeds0[k - 1], cws[k - 1], box["t"][k - 1], 0)
                extra[name] = {"value": world * lv / (msf * 1e-3), "unit": "leaves/s", "ms_per_step": msf, "in_bits": n,
                               "keys_per_gpu": k, "lsu_roofline_frac": lsu_frac(lv, 2.0, mskf),
                               "note": "includes the torch.empty of the output"}
            box.clear()
            del cws, s0s
            torch.cuda.empty_cache()
            # f-1 / f-2: DPF Gen, relayout, packed-row evaluation on the C2 shape
            kf = args.keys   # (whole waves matter: 2^22 keys = 36.9 waves of 113 664 resident keys; 2^20 = 9.2 -> 8 % tail)
            cf = fss_b200.Context("dpf", 32, "bytes", prg="aes128_mmo")
            s0s, alphas, betas, xs, cws = make_keys(cf, kf, gen)
            seeds0 = s0s[:, 0].contiguous()
            alf = cf.in_tensor(alphas, dev)
            msf, mskf = timed(lambda: cf.gen(s0s, alf, betas), x_steps, x_warm)
            extra["gen_dpf_n32_bytes"] = {"value": world * kf / (msf * 1e-3), "unit": "keys/s", "ms_per_step": msf, "keys_per_gpu": kf,
                                          "aes_blocks_per_key": 4 * 32, "lsu_roofline_frac": lsu_frac(kf, 4 * 32, mskf)}
            lay = cf.relayout(cws)
            msf, mskf = timed(lambda: cf.relayout(cws, out=lay), x_steps, x_warm)
            moved = cws.numel() * 4 + sum(t.numel() * 4 for t in lay if t is not None)
            extra["relayout_dpf_n32"] = {"value": world * kf / (msf * 1e-3), "unit": "keys/s", "ms_per_step": msf, "keys_per_gpu": kf,
                                         "hbm_gbs": moved / (mskf * 1e-3) / 1e9, "hbm_frac": moved / (mskf * 1e-3) / 1e9 / hbm_peak,
                                         "note": "bytes read + written over the kernel time, output arrays preallocated"}
            del lay
            kpk = kf
            prow = cf.pack_rows(cws[:kpk].cpu()).to(dev)
            yp = torch.empty((kpk, 4), dtype=torch.int32, device=dev)
            msf, mskf = timed(lambda: cf.eval_packed(0, seeds0[:kpk], prow, xs[:kpk], out=yp), x_steps, x_warm)
            if chk:
                chk.point("packed rows dpf n=32", cf, 0, seeds0[:kpk], cws[:kpk], xs[:kpk], yp)
            extra["eval_packed_dpf_n32"] = {"value": world * kpk / (msf * 1e-3), "unit": "evals/s", "ms_per_step": msf,
                                            "keys_per_gpu": kpk, "lsu_roofline_frac": lsu_frac(kpk, 32, mskf)}
            del prow, yp, cws, s0s, betas
            torch.cuda.empty_cache()
        # the reference's primary GPU PRG (ChaCha<2>, 20 rounds) on the C2 shape: bound = integer issue rate
        k6 = min(args.keys, 1 << 21)
        c6 = fss_b200.Context("dpf", 32, "bytes", prg="chacha")
        s0s, alphas, betas, xs, cws = make_keys(c6, k6, gen)
        seeds0 = s0s[:, 0].contiguous()
        ys = torch.empty((k6, 4), dtype=torch.int32, device=dev)
        ms6, msk6 = timed(lambda: c6.eval(0, seeds0, cws, xs, out=ys), x_steps, x_warm)
        if chk:
            chk.point("dpf n=32 bytes chacha", c6, 0, seeds0, cws, xs, ys)
        mixed = peaks.get("lop3_imad_mixed")
        extra["dpf_n32_chacha"] = {
            "value": world * k6 / (ms6 * 1e-3), "unit": "evals/s", "ms_per_step": ms6, "keys_per_gpu": k6,
            "issue_roofline_frac": (k6 * 32 * (976 + 12) / (msk6 * 1e-3) / mixed) if mixed else None,
            "note": "976 integer instructions per ChaCha20 block (SURVEY.md section 8d) against the measured mixed "
                    "LOP3+IMAD issue rate (ALU + FMA pipes)"}
        del s0s, betas, cws, ys, seeds0
        torch.cuda.empty_cache()
        # VDPF: DPF walk + Blake3 proof tail, n=32, 2^20 keys
        k7 = min(args.keys, 1 << 21)
        c7 = fss_b200.Context("vdpf", 32, "bytes", prg="aes128_mmo")
        s0s = rand_i32((k7, 2, 4), gen)
        betas = rand_i32((k7, 4), gen)
        s0s[:, :, 3] &= ~1
        betas[:, 3] &= ~1
        alphas, xs = rand_i32((k7,), gen), rand_i32((k7,), gen)
        xs[::16] = alphas[::16]
        vcws, vcs, vocws, status = c7.vdpf_gen(s0s, alphas, betas)
        seeds0, seeds1 = s0s[:, 0].contiguous(), s0s[:, 1].contiguous()
        vbox = {}
        ms7, msk7 = timed(lambda: vbox.__setitem__("r", c7.vdpf_eval(0, seeds0, vcws, vcs, vocws, xs)), x_steps, x_warm)
        y0v, pi0 = vbox["r"]
        y1v, pi1 = c7.vdpf_eval(1, seeds1, vcws, vcs, vocws, xs)
        ok = status == 0                                   # Gen's status 1 = resample the seeds (vdpf.cuh:160-169)
        hitv = ((xs == alphas) & ok).unsqueeze(1)
        if not torch.equal((y0v ^ y1v)[ok], torch.where(hitv, betas, torch.zeros_like(betas))[ok]) or \
                not torch.equal(pi0[ok], pi1[ok]):
            raise SystemExit("bench: VDPF reconstruction / proof equality failed on the timed batch")
        if chk:
            chk.done.append(f"vdpf n=32: reconstruction and per-point proof equality on the whole timed batch ({int(ok.sum())} keys)")
        extra["vdpf_n32_aes"] = {"value": world * k7 / (ms7 * 1e-3), "unit": "evals/s", "ms_per_step": ms7,
                                 "keys_per_gpu": k7, "lsu_roofline_frac_aes_only": lsu_frac(k7, 32, msk7),
                                 "note": "n AES blocks for the walk + 2 Blake3 compressions per evaluation; includes the "
                                         "torch.empty of the outputs"}
        # Vdpf::Gen on the same batch, and Vdpf::EvalAll (n = 16): the proof of a key chains its 2^n leaf hashes IN ORDER
        # (vdpf.cuh:313-341), so a key is one sequential Blake3 chain and only keys run in parallel
        msg7, mskg7 = timed(lambda: c7.vdpf_gen(s0s, alphas, betas), x_steps, x_warm)
        extra["gen_vdpf_n32"] = {"value": world * k7 / (msg7 * 1e-3), "unit": "keys/s", "ms_per_step": msg7, "keys_per_gpu": k7,
                                 "note": "4 AES blocks per level and key + the two parties' leaf hashes (4 Blake3 compressions)"}
        del vcws, vcs, vocws, vbox
        c8 = fss_b200.Context("vdpf", 16, "bytes", prg="aes128_mmo")
        k8 = 2048
        a8 = alphas[:k8] & 0xFFFF
        v8 = c8.vdpf_gen(s0s[:k8], a8, betas[:k8])
        abox = {}
        ms8, msk8 = timed(lambda: abox.__setitem__("r", c8.vdpf_eval_all(0, seeds0[:k8], v8[0], v8[1], v8[2])), max(2, x_steps // 4), 2)
        ya0, pa0 = abox["r"]
        ya1, pa1 = c8.vdpf_eval_all(1, seeds1[:k8], v8[0], v8[1], v8[2])
        ok8 = v8[3] == 0
        rec = (ya0 ^ ya1)[ok8]
        a8l = a8[ok8].to(torch.int64) & 0xFFFF
        idx = torch.arange(rec.shape[0], device=dev)
        if not torch.equal(pa0[ok8], pa1[ok8]) or not torch.equal(rec[idx, a8l], betas[:k8][ok8]) or \
                int((rec != 0).any(dim=2).sum()) != int((betas[:k8][ok8] != 0).any(dim=1).sum()):
            raise SystemExit("bench: VDPF EvalAll reconstruction / proof equality failed on the timed batch")
        if chk:
            chk.done.append(f"vdpf evalall n=16: both parties' proofs equal and shares reconstruct the point function ({int(ok8.sum())} keys)")
        extra["vdpf_evalall_n16"] = {"value": world * k8 * 65536 / (ms8 * 1e-3), "unit": "leaves/s", "ms_per_step": ms8, "keys_per_gpu": k8,
                                     "note": "tree + 2 Blake3 compressions per leaf in parallel, then one sequential hash chain of "
                                             "2^16 links per key (a warp per key): latency-bound by the chain's definition"}
        del s0s, betas, seeds0, seeds1, abox, ya0, ya1, v8
        torch.cuda.empty_cache()
        sampler.stop()

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        eng = cpu_engine()
        threads = host_threads()
        rate0, _ = cpu_dpf_eval_rate(eng, 1 << 14, threads, 2)
        sample = int(min(KEYS_PER_GPU, max(1 << 14, rate0 * 5.0)))
        sample = 1 << (sample.bit_length() - 1)
        rate, secs = cpu_dpf_eval_rate(eng, sample, threads, 3)
        rate1, _ = cpu_dpf_eval_rate(eng, 1 << 14, 1, 2)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": eng.kind,
                        "sample": f"{sample} of the 2^22 keys, best of 3 passes ({secs:.2f} s each), reference "
                                  f"Dpf::Eval with Aes128Mmo<2> (OpenSSL AES-NI), one EVP ctx set per thread",
                        "single_thread_value": rate1}
        try:
            cpu_baseline["others"] = cpu_other_rates(eng, threads)
        except Exception as e:   # (an engine without these instantiations: the headline baseline stands on its own)
            cpu_baseline["others"] = {"unavailable": str(e)}
        if eng.kind == "reference":   # (ii) of SURVEY.md section 8d: the same Eval with the reference's raw AES-NI PRG
            try:
                rate_raw, _ = cpu_dpf_eval_rate(eng, sample, threads, 2, prg="aes128_mmo_raw")
                cpu_baseline["aesni_raw_value"] = rate_raw
                cpu_baseline["aesni_raw_note"] = ("reference Dpf::Eval with Aes128MmoRaw<2> (AES-NI intrinsics, no EVP call "
                                                  "overhead), same sample and threads: the CPU ceiling, not the PRG north_star names")
            except Exception as e:   # an engine built without the raw PRG: the headline baseline stands on its own
                cpu_baseline["aesni_raw_value"] = None
                cpu_baseline["aesni_raw_note"] = f"unavailable: {e}"

    if rank == 0:
        roofline["kernels"] = extra   # (the driver keeps `roofline`; `extra` is kept for older readers)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "batched DPF Eval, n=32, 1 input per key, group::Bytes, AES-128 MMO "
                                   "(BASELINE configs[1])", "in_bits": N_BITS, "keys_per_gpu": nkeys,
                       "keys_total": world * nkeys, "party": 0, "cw_layout": "key-major Dpf::Cw (reference layout)",
                       "parallelism": f"keys sharded over {world} GPU(s), no collective",
                       "l2": "inputs (4.4 GB per GPU) are far larger than the 126 MB L2; no explicit flush",
                       "checked": (chk.done if chk else ["--no-check"]), "checker": (chk.kind if chk else None)},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": sampler.summary(), "strong_scaling": strong, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# one process, N GPUs: the multi-device entry points of the C ABI (no torchrun, no NCCL)
# ---------------------------------------------------------------------------------------------------------

def run_single_process(args) -> None:
    import torch

    from fss_b200.multi import MultiContext

    ndev = args.gpus
    if torch.cuda.device_count() < ndev:
        raise SystemExit(f"--single-process --gpus {ndev}: only {torch.cuda.device_count()} GPUs visible")
    devs = list(range(ndev))
    mc = MultiContext(devs, "dpf", N_BITS, "bytes", prg="aes128_mmo")
    nkeys = args.keys
    chk = Checker() if not args.no_check else None
    s0s, alphas, betas, xs = [], [], [], []
    for d in devs:
        with torch.cuda.device(d):
            g = torch.Generator(device=f"cuda:{d}").manual_seed(42 + d)
            r = lambda shape: torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int64, device=f"cuda:{d}",  # noqa: E731
                                            generator=g).to(torch.int32)
            s, b = r((nkeys, 2, 4)), r((nkeys, 4))
            s[:, :, 3] &= ~1
            b[:, 3] &= ~1
            a, x = r((nkeys,)), r((nkeys,))
            x[::16] = a[::16]
            s0s.append(s), betas.append(b), alphas.append(a), xs.append(x)
    cws = mc.gen(s0s, alphas, betas)
    seeds0 = [s[:, 0].contiguous() for s in s0s]
    ys = [torch.empty((nkeys, 4), dtype=torch.int32, device=f"cuda:{d}") for d in devs]
    mc.sync()
    for _ in range(args.warmup):
        mc.eval(0, seeds0, cws, xs, out=ys)
    mc.sync()
    # one event pair per device around the K launches on that device's stream; job time = max over devices
    ev = []
    for d in devs:
        with torch.cuda.device(d):
            ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            ev[-1][0].record(torch.cuda.current_stream(d))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mc.eval(0, seeds0, cws, xs, out=ys)
    for d in devs:
        with torch.cuda.device(d):
            ev[d][1].record(torch.cuda.current_stream(d))
    mc.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    ms_step = max(a.elapsed_time(b) for a, b in ev) / args.steps
    if chk:
        for d in devs:
            chk.point(f"single-process cuda:{d}", mc.ctxs[d], 0, seeds0[d], cws[d], xs[d], ys[d], count=1 << 10)
    # e2e: host arrays of the whole batch (ndev x 2^22 keys) through fssb200_eval_host_multi
    e2e = None
    if not args.no_e2e:
        h_seeds = torch.cat([t.cpu() for t in seeds0]).pin_memory()
        h_cws = torch.cat([t.cpu() for t in cws]).pin_memory()
        h_xs = torch.cat([t.cpu() for t in xs]).pin_memory()
        h_ys = torch.empty((ndev * nkeys, 4), dtype=torch.int32).pin_memory()
        want_ys = torch.cat([t.cpu() for t in ys])
        e_steps = 3

        def host_leg():
            for _ in range(2):
                mc.eval_host(0, h_seeds, h_cws, h_xs, out=h_ys)
            h_ys.zero_()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                mc.eval_host(0, h_seeds, h_cws, h_xs, out=h_ys)
            dt = (time.perf_counter() - t0) / e_steps
            if not torch.equal(h_ys, want_ys):
                raise SystemExit("bench: fssb200_eval_host_multi disagrees with the device path")
            return dt

        dt = host_leg()                                  # the library's choice (balanced blocks from 3 devices on)
        os.environ["FSSB200_MULTI_BALANCE"] = "0"
        dt_static = host_leg()                           # equal contiguous ranges per device (= what N processes do)
        del os.environ["FSSB200_MULTI_BALANCE"]
        e2e = {"value": ndev * nkeys / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": e_steps,
               "h2d_bytes_per_step": ndev * nkeys * (33 * 32 + 16 + 4), "d2h_bytes_per_step": ndev * nkeys * 16,
               "h2d_gbs": ndev * nkeys * (33 * 32 + 16 + 4) / dt / 1e9,
               "static_split": {"value": ndev * nkeys / dt_static, "ms_per_step": dt_static * 1e3,
                                "note": "FSSB200_MULTI_BALANCE=0: device d takes key_shard(d); lasts as long as the slowest link"},
               "api": "fssb200_eval_host_multi: one process, host arrays of the whole batch; from 3 devices on the devices "
                      "claim key blocks from one counter (two calls in flight per device)"}
    line = {
        "metric": METRIC, "value": ndev * nkeys / (ms_step * 1e-3), "unit": UNIT, "n_gpus": ndev, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "wall_ms_per_step": wall_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "batched DPF Eval, n=32, 1 input per key, group::Bytes, AES-128 MMO (BASELINE configs[1])",
                   "in_bits": N_BITS, "keys_per_gpu": nkeys, "keys_total": ndev * nkeys,
                   "parallelism": f"ONE process, {ndev} GPUs, fssb200_eval_multi (one stream + event pair per device), "
                                  "no collective", "checked": (chk.done if chk else ["--no-check"])},
        "e2e": e2e, "gpu_launches": args.steps * ndev,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--keys", type=int, default=KEYS_PER_GPU, help="keys per GPU (default 2^22)")
    ap.add_argument("--evalall-bits", type=int, default=28)
    ap.add_argument("--evalall-keys", type=int, default=8)
    ap.add_argument("--single-process", action="store_true", help="one process drives --gpus GPUs (fssb200_*_multi)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-frows", action="store_true", help="skip the SURVEY section 8(f) rows of `extra`")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle checks of the timed outputs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
