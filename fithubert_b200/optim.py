"""Fused multi-tensor AdamW (K11) + the s3prl warm-up / linear-decay schedule, and the data-parallel
gradient all-reduce (K12).  Replaces s3prl get_optimizer (reference train.py:407-421) and Lightning's
DDP reducer (train.py:494)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import kernels as K
from . import lib as L


def warmup_linear(step: int, total_steps: int, warmup: float) -> float:
    """s3prl 'warmup_linear': x/w for x < w else max((x - 1)/(w - 1), 0), x = step/total."""
    x = step / max(1, total_steps)
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0.0)


class FusedAdamW:
    """One kernel launch per optimizer step over every student parameter.  Gradients are read from the
    engine's flat gradient buffer through per-tensor 3-D strides (no layout conversion pass)."""

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=1e-6, total_steps=0,
                 warmup_proportion=0.05, mode="s3prl"):
        self.model = model
        self.lr, self.betas, self.eps, self.wd = float(lr), tuple(betas), float(eps), float(weight_decay)
        self.total_steps, self.warmup, self.mode = int(total_steps), float(warmup_proportion), mode
        self.step_count = 0
        self._table = None

    def _build(self):
        P, W, G = self.model.engine_state(True)
        dev = G.flat.device
        layout = tuple((pn, e[0], e[1]) for pn, e in G.entries.items())
        if getattr(self, "m", None) is not None and self.m.device == dev and layout == getattr(self, "_layout", None):
            pass  # same gradient layout (a rebuilt buffer, moved parameters): the moments stay
        elif self.step_count > 0 and getattr(self, "m", None) is not None:
            raise L.FhbError("the student's parameter set changed after optimizer steps were taken: Adam moments cannot "
                             "be carried over (build a new optimizer, or restore one from a checkpoint)")
        else:
            self.m = torch.zeros(G.numel, device=dev, dtype=torch.float32)
            self.v = torch.zeros(G.numel, device=dev, dtype=torch.float32)
        self._layout = layout
        self.P, self.W, self.G = P, W, G
        entries, mx = [], 1
        for pn, (off, n, dims, gs) in G.entries.items():
            e = L.AdamwTensor()
            e.p = P[pn].data_ptr()
            e.g = G.flat.data_ptr() + 4 * off
            e.m = self.m.data_ptr() + 4 * off
            e.v = self.v.data_ptr() + 4 * off
            e.n = n
            e.dim = (L.C.c_int64 * 3)(*dims)
            e.gstride = (L.C.c_int64 * 3)(*gs)
            entries.append(e)
            mx = max(mx, n)
        self._table = L.table_to_device(entries, dev)
        self._n, self._max_n = len(entries), mx
        self._ptrs = tuple(P[pn].data_ptr() for pn in G.entries)

    def current_lr(self) -> float:
        if self.total_steps <= 0:
            return self.lr
        return self.lr * warmup_linear(self.step_count, self.total_steps, self.warmup)

    def step(self, grad_scale: float = 1.0):
        P, W, G = self.model.engine_state(True)
        if self._table is None or G is not self.G or self._ptrs != tuple(P[pn].data_ptr() for pn in G.entries):
            self._build()
        self.W = W
        self.step_count += 1
        b1, b2 = self.betas
        K.adamw_multi(self._table, self._n, self._max_n, self.current_lr(), b1, b2, self.eps, self.wd,
                      self.step_count, 0 if self.mode == "s3prl" else 1, grad_scale)
        self.W.mark_stale()  # parameters changed behind torch's back: re-derive the bf16 shadows

    def zero_grad(self):
        _, _, G = self.model.engine_state(True)
        G.zero_()


class GradAllReduce:
    """Data-parallel gradient exchange: bucketed NCCL all-reduce of the flat gradient buffer on a side
    stream; buckets are launched as backward finishes the corresponding segment (reverse-forward order:
    heads -> layers -> front-end), the optimizer waits on the last bucket.  Averaging (1/world) is folded
    into the AdamW kernel's grad_scale."""

    def __init__(self, n_buckets: int = 0):
        if n_buckets <= 0:
            import os
            n_buckets = int(os.environ.get("FHB_REDUCE_BUCKETS", "6"))
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.enabled else 1
        self.n_buckets = n_buckets
        self.stream = torch.cuda.Stream() if (self.enabled and torch.cuda.is_available()) else None
        self._handles = []

    def bucket_bounds(self, numel: int, start: int = 0):
        """Bucket boundaries over flat[start:numel]; bucket size = ceil(buffer size / n_buckets) whatever the range."""
        step = -(-max(numel, getattr(self, "_full", numel)) // self.n_buckets)
        step = (step + 3) // 4 * 4
        return [(a, min(a + step, numel)) for a in range(start, numel, step)]

    def _issue(self, flat: torch.Tensor, lo: int, hi: int):
        bounds = [(a, b) for (a, b) in self.bucket_bounds(hi, lo)]
        if self.stream is None:  # gloo / CPU tests
            for (a, b) in reversed(bounds):
                dist.all_reduce(flat[a:b])
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for (a, b) in reversed(bounds):
                dist.all_reduce(flat[a:b])

    def reduce_tail(self, flat: torch.Tensor, start: int):
        """flat[start:] is final (backward fills the buffer from its end: heads, then layers 12..1, then the front end):
        reduce the part not sent yet, flat[start : previous start], now - under the rest of the backward.  reduce_all()
        sends what is left."""
        if not self.enabled:
            return
        start = start // 4 * 4
        hi = getattr(self, "_tail_from", None)
        hi = flat.numel() if hi is None else hi
        if start >= hi:
            return
        self._issue(flat, start, hi)
        self._tail_from = start

    def reduce_all(self, flat: torch.Tensor):
        """Every bucket not sent yet, issued back to front (the order backward completes them)."""
        if not self.enabled:
            return
        hi = getattr(self, "_tail_from", None)
        self._tail_from = None
        if hi is None or hi > 0:
            self._issue(flat, 0, flat.numel() if hi is None else hi)

    def wait(self):
        if self.enabled and self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
