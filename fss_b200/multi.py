"""One process, several GPUs: torch-facing wrapper of the multi-device entry points of the C ABI
(include/fssb200.h: fssb200_eval_multi / fssb200_eval_all_multi / fssb200_gen_multi / fssb200_eval_host_multi).

The path shards with no exchange step (keys are independent, dpf.cuh:170-214; an EvalAll subtree depends only on
its root, dpf.cuh:291-301): a multi-device call is one stream-ordered launch per device on that device's tensors.
Nothing is copied between devices and no collective runs.  The process-per-GPU launcher (torchrun + NCCL barrier)
lives in ``fss_b200.sharding`` / ``bench.py``; this module is the single-process alternative SURVEY.md section 8e
describes ("one process, one stream + events per device").
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .context import Context, IntLike


def key_shard(nkeys: int, d: int, n: int) -> Tuple[int, int]:
    """fssb200_key_shard: [begin, end) of shard d of n."""
    b, e = C.c_size_t(0), C.c_size_t(0)
    L.check(L.lib.fssb200_key_shard(nkeys, d, n, C.byref(b), C.byref(e)), "fssb200_key_shard")
    return int(b.value), int(e.value)


class MultiContext:
    """The same parameter set on several devices (``devices``: CUDA ordinals; an ordinal may repeat, which runs several
    shards on one GPU -- used by the single-GPU tests)."""

    def __init__(self, devices: Sequence[int], scheme: str, in_bits: int, group: str = "bytes", **kw):
        self.devices = list(devices)
        self.ctxs = [Context(scheme, in_bits, group, **kw) for _ in self.devices]
        self.handles = [c.handle(d) for c, d in zip(self.ctxs, self.devices)]
        self.ndev = len(self.devices)
        self._harr = (C.c_void_p * self.ndev)(*[h.value for h in self.handles])
        self.scheme, self.in_bits = scheme, in_bits
        self.in_bytes, self.ncw = self.ctxs[0].in_bytes, self.ctxs[0].ncw

    def close(self) -> None:
        for c in self.ctxs:
            c.close()

    def set_host_mode(self, mode: int) -> None:
        """fssb200_ctx_set_host_mode on every device's context (eval_host)."""
        for c, d in zip(self.ctxs, self.devices):
            c.set_host_mode(mode, d)

    def launch_count(self) -> int:
        return sum(c.launch_count(d) for c, d in zip(self.ctxs, self.devices))

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _ptrs(self, ts: Optional[Sequence[Optional[torch.Tensor]]]):
        if ts is None:
            return None
        for t, d in zip(ts, self.devices):
            if t is not None and (t.device.type != "cuda" or t.device.index != d):
                raise RuntimeError(f"per-device tensor lives on {t.device}, expected cuda:{d}")
        return (C.c_void_p * self.ndev)(*[None if t is None else t.data_ptr() for t in ts])

    def _streams(self):
        return (C.c_void_p * self.ndev)(*[torch.cuda.current_stream(d).cuda_stream for d in self.devices])

    def _check(self, rc: int, rcs, what: str) -> None:
        if rc != 0:
            per_dev = ", ".join(f"cuda:{d}={L.strerror(int(r))}" for d, r in zip(self.devices, rcs) if r)
            raise L.FssError(rc, f"{what} [{per_dev}]")

    def leaf_shard(self, d: int) -> Tuple[int, int]:
        """fssb200_leaf_shard: (leaf_begin, leaf_count) of shard d; count 0 = empty shard."""
        b, c = C.c_uint64(0), C.c_uint64(0)
        L.check(L.lib.fssb200_leaf_shard(self.handles[d], d, self.ndev, C.byref(b), C.byref(c)), "fssb200_leaf_shard")
        return int(b.value), int(c.value)

    # ---- device tensors, one per device ---------------------------------------------------------------------------
    def gen(self, s0s: Sequence[torch.Tensor], alphas: Sequence[IntLike], betas: Optional[Sequence[torch.Tensor]]):
        s0s = [t.contiguous() for t in s0s]
        al = [c.in_tensor(a, t.device) for c, a, t in zip(self.ctxs, alphas, s0s)]
        be = None if betas is None else [t.contiguous() for t in betas]
        n = [t.shape[0] for t in s0s]
        cws = [torch.empty((k, self.ncw, 8), dtype=torch.int32, device=t.device) for k, t in zip(n, s0s)]
        ocws = [torch.empty((k, 4), dtype=torch.int32, device=t.device) for k, t in zip(n, s0s)] \
            if self.scheme == "halftree" else None
        rcs = (C.c_int * self.ndev)()
        rc = L.lib.fssb200_gen_multi(self._harr, self.ndev, self._ptrs(s0s), self._ptrs(al), self._ptrs(be),
                                     self._ptrs(cws), self._ptrs(ocws), (C.c_size_t * self.ndev)(*n), self._streams(),
                                     rcs)
        self._check(rc, rcs, "fssb200_gen_multi")
        return (cws, ocws) if ocws is not None else cws

    def eval(self, party: int, seeds: Sequence[torch.Tensor], cws: Sequence[torch.Tensor], xs: Sequence[IntLike],
             ocws: Optional[Sequence[torch.Tensor]] = None, out: Optional[Sequence[torch.Tensor]] = None
             ) -> List[torch.Tensor]:
        seeds, cws = [t.contiguous() for t in seeds], [t.contiguous() for t in cws]
        x = [c.in_tensor(v, t.device) for c, v, t in zip(self.ctxs, xs, seeds)]
        n = [t.shape[0] for t in seeds]
        for k, a, b in zip(n, cws, x):
            if a.shape[0] != k or b.shape[0] != k:
                raise TypeError("per-device cws / xs must have as many rows as seeds")
        ys = list(out) if out is not None else [torch.empty((k, 4), dtype=torch.int32, device=t.device)
                                                for k, t in zip(n, seeds)]
        rcs = (C.c_int * self.ndev)()
        rc = L.lib.fssb200_eval_multi(self._harr, self.ndev, party, self._ptrs(seeds), self._ptrs(cws),
                                      self._ptrs(ocws), self._ptrs(x), self._ptrs(ys), (C.c_size_t * self.ndev)(*n),
                                      self._streams(), rcs)
        self._check(rc, rcs, "fssb200_eval_multi")
        return ys

    def eval_all(self, party: int, seeds: Sequence[torch.Tensor], cws: Sequence[torch.Tensor],
                 ocws: Optional[Sequence[torch.Tensor]] = None, leaf_ranges: Optional[Sequence[Tuple[int, int]]] = None,
                 out: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
        """leaf_ranges[d] = (leaf_begin, leaf_count) of device d (default: the whole domain on every device)."""
        seeds, cws = [t.contiguous() for t in seeds], [t.contiguous() for t in cws]
        n = [t.shape[0] for t in seeds]
        full = 1 << self.in_bits
        lr = list(leaf_ranges) if leaf_ranges is not None else [(0, full)] * self.ndev
        if any(c <= 0 for _, c in lr):
            raise ValueError("empty leaf range: drop that device from the call")
        if out is not None:
            ys = list(out)
        elif self.scheme == "grotto":
            ys = [torch.empty((k, c), dtype=torch.uint8, device=t.device) for k, (_, c), t in zip(n, lr, seeds)]
        else:
            ys = [torch.empty((k, c, 4), dtype=torch.int32, device=t.device) for k, (_, c), t in zip(n, lr, seeds)]
        rcs = (C.c_int * self.ndev)()
        rc = L.lib.fssb200_eval_all_multi(self._harr, self.ndev, party, self._ptrs(seeds), self._ptrs(cws),
                                          self._ptrs(ocws), self._ptrs(ys), (C.c_size_t * self.ndev)(*n),
                                          (C.c_uint64 * self.ndev)(*[b for b, _ in lr]),
                                          (C.c_uint64 * self.ndev)(*[c for _, c in lr]), self._streams(), rcs)
        self._check(rc, rcs, "fssb200_eval_all_multi")
        return ys

    def sync(self) -> None:
        rcs = (C.c_int * self.ndev)()
        rc = L.lib.fssb200_multi_sync(self._harr, self.ndev, self._streams(), rcs)
        self._check(rc, rcs, "fssb200_multi_sync")

    # ---- host tensors of the whole batch ------------------------------------------------------------------------------
    def eval_host(self, party: int, seeds: torch.Tensor, cws: torch.Tensor, xs: IntLike,
                  ocws: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fssb200_eval_host_multi: CPU tensors of the whole batch.  Device d evaluates key_shard(N, d, ndev), or -- when
        the keys cross in the reference layout (>= 3 devices, or host mode 1) -- the devices claim key blocks from one
        counter so that unequal links finish together."""
        if seeds.device.type != "cpu":
            raise RuntimeError("eval_host takes CPU tensors")
        seeds, cws = seeds.contiguous(), cws.contiguous()
        n = seeds.shape[0]
        x = self.ctxs[0].in_tensor(xs, seeds.device)
        if cws.shape[0] != n or x.shape[0] != n:
            raise TypeError(f"cws / xs must have {n} rows")
        ocws = None if ocws is None else ocws.contiguous()
        ys = out if out is not None else torch.empty((n, 4), dtype=torch.int32)
        rcs = (C.c_int * self.ndev)()
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        rc = L.lib.fssb200_eval_host_multi(self._harr, self.ndev, party, p(seeds), p(cws), p(ocws), p(x), p(ys), n, rcs)
        self._check(rc, rcs, "fssb200_eval_host_multi")
        return ys
