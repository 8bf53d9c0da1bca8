#!/usr/bin/env python
"""bench.py -- throughput of the DPF/DCF PRG-tree hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] -- batched DPF Eval, n = 32, one input per key,
group::Bytes, AES-128 MMO PRG, 2^22 independent keys PER GPU (weak scaling: keys are independent, every
rank evaluates its own key range, no data-path collective).  A step = one pass of the hot path over the
rank's batch = one launch of the point-evaluation kernel.  Inputs are synthetic (torch RNG on the device,
keys produced by this library's own Gen kernel) and 4.4 GB per GPU, i.e. far larger than the 126 MB L2.

The one JSON line also carries: the other BASELINE configs as `extra` (DCF n=64 u127, DPF EvalAll n=28,
Half-Tree n=32), `e2e` (the same metric through the host-buffer C-ABI entry point with pinned host
buffers, copies inside the timed region), `roofline` (integer-pipe roofline of SURVEY.md section 8d against
an on-box LOP3 issue-rate measurement, plus HBM numbers), `cpu_baseline` (the reference's own Eval with its
OpenSSL AES-NI PRG on the host cores, rank 0, N=1 only) and `clocks`.

`--impl reference` times the reference's CPU implementation of the same path (oracle/_ref, built from the
unmodified reference headers; falls back to the plain-C port) on all host threads.
"""
from __future__ import annotations

import argparse
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


def cpu_dpf_eval_rate(engine, nkeys: int, threads: int, repeats: int = 1, cws=None, seeds=None, xs=None):
    """Reference Dpf::Eval (dpf.cuh:170-214, Aes128Mmo<2> = OpenSSL AES-NI) over nkeys keys on `threads` host
    threads, one PRG context set per thread.  Returns (evals/s, seconds per pass)."""
    import numpy as np
    from oracle import Params, synth_inputs
    p = Params(scheme="dpf", in_bits=N_BITS, group="bytes", prg="aes128_mmo")
    if cws is None:
        s0s, alphas, betas, xs = synth_inputs(p, nkeys, seed=42)
        cws = engine.gen(p, s0s, alphas, betas, threads=threads)
        seeds = np.ascontiguousarray(s0s[:, 0])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        engine.eval(p, 0, seeds, cws, xs, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return nkeys / best, best


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import Params, synth_inputs
    eng = cpu_engine()
    threads = host_threads()
    p = Params(scheme="dpf", in_bits=N_BITS, group="bytes", prg="aes128_mmo")
    # size the per-step sample so that warmup + steps stay within a few minutes: calibrate on 2^14 keys
    s0s, alphas, betas, xs = synth_inputs(p, 1 << 14, seed=7)
    cws = eng.gen(p, s0s, alphas, betas, threads=threads)
    rate, _ = cpu_dpf_eval_rate(eng, 1 << 14, threads, 2, cws, np.ascontiguousarray(s0s[:, 0]), xs)
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    sample = int(min(KEYS_PER_GPU, max(1 << 14, rate * min(budget_s, 2.0))))
    sample = 1 << (sample.bit_length() - 1)
    s0s, alphas, betas, xs = synth_inputs(p, sample, seed=42)
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
        "config": {"workload": "batched DPF Eval, n=32, group::Bytes, Aes128Mmo<2> (OpenSSL AES-NI), reference "
                               "CPU path", "in_bits": N_BITS, "keys_per_step": sample, "sample": sample_desc},
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
# own arm
# ---------------------------------------------------------------------------------------------------------

def run_own_arm(args) -> None:
    import torch
    import torch.distributed as dist

    import fss_b200  # fails loudly if libfssb200.so is missing

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

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rand_i32(shape, gen):
        return torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int64, device=dev, generator=gen).to(torch.int32)

    def make_keys(ctx, nkeys, gen, wide=False):
        s0s = rand_i32((nkeys, 2, 4), gen)
        betas = rand_i32((nkeys, 4), gen)
        s0s[:, :, 3] &= ~1
        betas[:, 3] &= ~1
        if wide:
            alphas = torch.randint(-2 ** 63, 2 ** 63 - 1, (nkeys,), dtype=torch.int64, device=dev, generator=gen)
            xs = torch.randint(-2 ** 63, 2 ** 63 - 1, (nkeys,), dtype=torch.int64, device=dev, generator=gen)
        else:
            alphas, xs = rand_i32((nkeys,), gen), rand_i32((nkeys,), gen)
        xs[::16] = alphas[::16]  # exercise the beta branch
        r = ctx.gen(s0s, alphas, betas)
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
        return max_over_ranks(total_ms / steps), per_launch

    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    sampler = ClockSampler(local)
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
    # correctness guard inside the bench: reconstruction on the timed batch
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
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("dpf_point_c2")
    except Exception:
        pass
    int_achieved = nkeys * OPS_PER_EVAL_C2 / (ms_kernel * 1e-3) / 1e12
    int_peak = (peaks.get("lop3", 0.0) / 1e12) or None
    lds_achieved = nkeys * N_BITS * LOOKUPS_PER_AES / (ms_kernel * 1e-3) / 1e12
    lds_peak = (peaks.get("lds32_conflict_free", 0.0) / 1e12) or None
    hbm_achieved = nkeys * BYTES_PER_EVAL_C2 / (ms_kernel * 1e-3) / 1e9
    # The binding roofline of a T-table AES kernel is the shared-memory lookup pipe (32 conflict-free LDS.32 lanes per
    # clock and SM, measured on this box in this run); the integer-ALU roofline of SURVEY.md section 8d (canonical 444
    # ops per block against the LOP3 issue rate) and the HBM figures are reported beside it.  Not HBM-bound: see `hbm`.
    roofline = {
        "bound": "smem_lsu", "achieved": lds_achieved, "peak": lds_peak, "unit": "Tlookups/s",
        "frac": (lds_achieved / lds_peak) if lds_peak else None, "traffic": traffic,
        "kernel": "point_kernel<DPF,Bytes,AES>", "ms_per_launch": ms_kernel,
        "algorithmic_lookups_per_eval": N_BITS * LOOKUPS_PER_AES,
        "peak_source": "on-box conflict-free ld.shared.u32 rate measured in this run (fssb200_microbench kind 3); "
                       "MEASURED_PEAKS.json has no shared-memory or integer peak",
        "int_alu": {"bound": "int_alu", "achieved": int_achieved, "peak": int_peak, "unit": "Tops/s(int32)",
                    "frac": (int_achieved / int_peak) if int_peak else None,
                    "algorithmic_ops_per_eval": OPS_PER_EVAL_C2,
                    "note": "canonical T-table count of SURVEY.md section 8d (444 ops per AES block + 12 glue per level) "
                            "against the measured LOP3 issue rate; this implementation issues ~240 ALU-pipe + ~90 "
                            "FMA-pipe integer instructions per level, so the fraction can exceed 1"},
        "microbench_ops_per_s": peaks,
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_achieved / hbm_peak, "algorithmic_bytes_per_eval": BYTES_PER_EVAL_C2,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if measured else "fallback 6.65 TB/s"},
        "aes_blocks_per_s": nkeys * 32 / (ms_kernel * 1e-3),
    }

    # ---- e2e: host buffers through the C ABI -------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_seeds = seeds0.cpu().pin_memory()
        h_cws = cws.cpu().pin_memory()
        h_xs = xs.cpu().pin_memory()
        h_ys = torch.empty((nkeys, 4), dtype=torch.int32).pin_memory()
        ctx.reserve_host(1 << 18)
        e_steps = max(1, min(args.steps, 5))
        for _ in range(2):
            ctx.eval(0, h_seeds, h_cws, h_xs, out=h_ys)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.eval(0, h_seeds, h_cws, h_xs, out=h_ys)   # returns when ys is complete on the host
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / e_steps)
        if not torch.equal(h_ys, ys.cpu()):
            raise SystemExit("bench: host-buffer path disagrees with the device path")
        pack_threads = ctx.host_pack_threads(local)
        cws_bytes = nkeys * ctx.packed_row_bytes(local) if pack_threads else h_cws.numel() * 4
        e2e = {"value": world * nkeys / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": e_steps,
               "h2d_bytes_per_step": h_seeds.numel() * 4 + cws_bytes + h_xs.numel() * 4,
               "d2h_bytes_per_step": h_ys.numel() * 4,
               "host_pack_threads": pack_threads,
               "api": "fssb200_eval_host (reference-layout keys in pinned host buffers, 2^18-key chunks, 2 streams), wall "
                      "clock around the blocking call" + (
                          f"; {pack_threads} host threads strip the 15 padding bytes of each 32-byte Dpf::Cw into pinned "
                          "staging while the previous chunk is in flight, so 17 B per level cross PCIe" if pack_threads
                          else "; keys copied as they are (not enough host cores per rank to pack faster than PCIe)")}
        if pack_threads:
            # the same call with packing switched off: the reference layout crosses PCIe as it is
            os.environ["FSSB200_PACK_THREADS"] = "0"
            ctx_d = fss_b200.Context("dpf", N_BITS, "bytes", prg="aes128_mmo")
            ctx_d.reserve_host(1 << 18, local)
            del os.environ["FSSB200_PACK_THREADS"]
            for _ in range(2):
                ctx_d.eval(0, h_seeds, h_cws, h_xs, out=h_ys)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                ctx_d.eval(0, h_seeds, h_cws, h_xs, out=h_ys)
            torch.cuda.synchronize()
            dt_d = max_over_ranks((time.perf_counter() - t0) / e_steps)
            e2e["direct_copy"] = {"value": world * nkeys / dt_d, "unit": UNIT, "ms_per_step": dt_d * 1e3,
                                  "h2d_bytes_per_step": h_seeds.numel() * 4 + h_cws.numel() * 4 + h_xs.numel() * 4}
            ctx_d.close()
        # the same keys in the compact level-major layout (fssb200_relayout: 16 B + 1 bit per level instead of the
        # 32-byte Dpf::Cw) through fssb200_eval_levelmajor_host
        lay = ctx.relayout(cws)
        torch.cuda.synchronize()
        ms_lm, _ = timed(lambda: ctx.eval_levelmajor(0, seeds0, lay, xs, out=ys), max(3, min(args.steps, 10)), 3)
        h_lay = tuple(None if t is None else t.cpu().pin_memory() for t in lay)
        for _ in range(2):
            ctx.eval_levelmajor(0, h_seeds, h_lay, h_xs, out=h_ys)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.eval_levelmajor(0, h_seeds, h_lay, h_xs, out=h_ys)
        torch.cuda.synchronize()
        dt_lm = max_over_ranks((time.perf_counter() - t0) / e_steps)
        if not torch.equal(h_ys, ys.cpu()):
            raise SystemExit("bench: level-major host path disagrees with the device path")
        e2e["compact_levelmajor"] = {
            "value": world * nkeys / dt_lm, "unit": UNIT, "ms_per_step": dt_lm * 1e3,
            "h2d_bytes_per_step": h_seeds.numel() * 4 + h_xs.numel() * 4 + sum(t.numel() * 4 for t in h_lay if t is not None),
            "d2h_bytes_per_step": h_ys.numel() * 4,
            "kernel_only_evals_per_s": world * nkeys / (ms_lm * 1e-3),
            "api": "fssb200_eval_levelmajor_host: keys held by the caller in the level-major layout of fssb200_relayout "
                   "(not the reference's Cw layout; reported beside the drop-in number, not instead of it)"}
        del h_seeds, h_cws, h_xs, h_ys, h_lay, lay
    del cws, s0s, betas, ys
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs (extra) --------------------------------------------------------------------------
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
        extra["dcf_n64_u127_aes"] = {
            "value": world * k3 / (ms3 * 1e-3), "unit": "evals/s", "ms_per_step": ms3, "keys_per_gpu": k3,
            "int_roofline_frac": (k3 * OPS_PER_EVAL_C3 / (msk3 * 1e-3) / 1e12 / int_peak) if int_peak else None,
            "lsu_roofline_frac": (k3 * 128 * LOOKUPS_PER_AES / (msk3 * 1e-3) / 1e12 / lds_peak) if lds_peak else None}
        del s0s, betas, cws, ys, seeds0
        torch.cuda.empty_cache()
        # C5: Half-Tree DPF n=32, 2^20 keys
        k5 = min(args.keys, 1 << 20)
        c5 = fss_b200.Context("halftree", 32, "bytes", prg="aes128_mmo")
        s0s, alphas, betas, xs, (cws, ocws) = make_keys(c5, k5, gen)
        seeds0 = s0s[:, 0].contiguous()
        ys = torch.empty((k5, 4), dtype=torch.int32, device=dev)
        ms5, msk5 = timed(lambda: c5.eval(0, seeds0, cws, xs, ocws, out=ys), x_steps, x_warm)
        extra["halftree_n32_aes"] = {
            "value": world * k5 / (ms5 * 1e-3), "unit": "evals/s", "ms_per_step": ms5, "keys_per_gpu": k5,
            "lsu_roofline_frac": (k5 * 32 * LOOKUPS_PER_AES / (msk5 * 1e-3) / 1e12 / lds_peak) if lds_peak else None}
        del s0s, betas, cws, ocws, ys, seeds0
        torch.cuda.empty_cache()
        # C4: DPF EvalAll n=28, 64 keys over 8 GPUs = 8 keys (32 GiB of leaves) per GPU and step
        n4, k4 = args.evalall_bits, args.evalall_keys
        c4 = fss_b200.Context("dpf", n4, "bytes", prg="aes128_mmo")
        s0s, alphas, betas, xs, cws = make_keys(c4, k4, gen)
        seeds0 = s0s[:, 0].contiguous()
        out = torch.empty((k4, 1 << n4, 4), dtype=torch.int32, device=dev)
        ms4, msk4 = timed(lambda: c4.eval_all(0, seeds0, cws, out=out), max(2, x_steps // 2), 2)
        leaves = k4 * (1 << n4)
        extra["dpf_evalall"] = {
            "value": world * leaves / (ms4 * 1e-3), "unit": "leaves/s", "ms_per_step": ms4, "in_bits": n4,
            "keys_per_gpu": k4, "output_gib_per_gpu": leaves * 16 / 2 ** 30,
            "int_roofline_frac": (leaves * OPS_PER_LEAF_C4 / (msk4 * 1e-3) / 1e12 / int_peak) if int_peak else None,
            "lsu_roofline_frac": (leaves * 2 * LOOKUPS_PER_AES / (msk4 * 1e-3) / 1e12 / lds_peak) if lds_peak else None,
            "hbm_write_gbs": leaves * 16 / (msk4 * 1e-3) / 1e9}
        # C4 as BASELINE configs[3] words it: "subtrees sharded" -- every rank expands ITS leaf range of every key
        # (8 keys per GPU in the job, so the per-GPU work stays 32 GiB of leaves: weak scaling), no collective
        if world > 1:
            from fss_b200.sharding import leaf_shard
            kk = k4 * world
            s0s, alphas, betas, xs, cws = make_keys(c4, kk, torch.Generator(device=dev).manual_seed(4242))  # same keys on every rank
            seeds0 = s0s[:, 0].contiguous()
            lb, lc = leaf_shard(n4, c4.granule(), rank, world)
            outv = out.view(-1)[: kk * lc * 4].view(kk, lc, 4)
            ms4s, _ = timed(lambda: c4.eval_all(0, seeds0, cws, leaf_begin=lb, leaf_count=lc, out=outv), max(2, x_steps // 2), 2)
            extra["dpf_evalall_subtree_sharded"] = {
                "value": kk * (1 << n4) / (ms4s * 1e-3), "unit": "leaves/s", "ms_per_step": ms4s, "in_bits": n4,
                "keys_total": kk, "leaf_range_per_gpu": [int(lb), int(lc)], "output_gib_per_gpu": kk * lc * 16 / 2 ** 30}
            del outv
        del out, cws
        torch.cuda.empty_cache()
        # the reference's primary GPU PRG (ChaCha<2>, 20 rounds) on the C2 shape: bound = integer issue rate
        k6 = min(args.keys, 1 << 21)
        c6 = fss_b200.Context("dpf", 32, "bytes", prg="chacha")
        s0s, alphas, betas, xs, cws = make_keys(c6, k6, gen)
        seeds0 = s0s[:, 0].contiguous()
        ys = torch.empty((k6, 4), dtype=torch.int32, device=dev)
        ms6, msk6 = timed(lambda: c6.eval(0, seeds0, cws, xs, out=ys), x_steps, x_warm)
        mixed = peaks.get("lop3_imad_mixed")
        extra["dpf_n32_chacha"] = {
            "value": world * k6 / (ms6 * 1e-3), "unit": "evals/s", "ms_per_step": ms6, "keys_per_gpu": k6,
            "issue_roofline_frac": (k6 * 32 * (976 + 12) / (msk6 * 1e-3) / mixed) if mixed else None,
            "note": "976 integer instructions per ChaCha20 block (SURVEY.md section 8d) against the measured mixed "
                    "LOP3+IMAD issue rate (ALU + FMA pipes)"}
        del s0s, betas, cws, ys, seeds0
        torch.cuda.empty_cache()
        # VDPF: DPF walk + Blake3 proof tail, n=32, 2^20 keys
        k7 = min(args.keys, 1 << 20)
        c7 = fss_b200.Context("vdpf", 32, "bytes", prg="aes128_mmo")
        s0s = rand_i32((k7, 2, 4), gen)
        betas = rand_i32((k7, 4), gen)
        s0s[:, :, 3] &= ~1
        betas[:, 3] &= ~1
        alphas, xs = rand_i32((k7,), gen), rand_i32((k7,), gen)
        vcws, vcs, vocws, _status = c7.vdpf_gen(s0s, alphas, betas)
        seeds0 = s0s[:, 0].contiguous()
        ms7, _ = timed(lambda: c7.vdpf_eval(0, seeds0, vcws, vcs, vocws, xs), x_steps, x_warm)
        extra["vdpf_n32_aes"] = {"value": world * k7 / (ms7 * 1e-3), "unit": "evals/s", "ms_per_step": ms7,
                                 "keys_per_gpu": k7}
        del s0s, betas, vcws, vcs, vocws, seeds0
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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "batched DPF Eval, n=32, 1 input per key, group::Bytes, AES-128 MMO "
                                   "(BASELINE configs[1])", "in_bits": N_BITS, "keys_per_gpu": nkeys,
                       "keys_total": world * nkeys, "party": 0, "cw_layout": "key-major Dpf::Cw (reference layout)",
                       "parallelism": f"keys sharded over {world} GPU(s), no collective",
                       "l2": "inputs (4.4 GB per GPU) are far larger than the 126 MB L2; no explicit flush"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": sampler.summary(), "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--keys", type=int, default=KEYS_PER_GPU, help="keys per GPU (default 2^22)")
    ap.add_argument("--evalall-bits", type=int, default=28)
    ap.add_argument("--evalall-keys", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
