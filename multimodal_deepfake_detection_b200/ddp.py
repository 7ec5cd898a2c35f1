"""Data parallelism for the clip path: one process per GPU, clips sharded by rank, gradients averaged with
bucketed all-reduces on a side stream that overlap the rest of backward.

The reference only has nn.DataParallel around LSTM+head (train_audio.py:16-18); the backbone is never parallel
there.  SURVEY.md §8(e): clips are independent through backbone, LSTM and head, so the only exchange step is
the gradient average.  BatchNorm statistics stay per-rank (the reference uses plain BatchNorm2d, no SyncBN).

How the overlap works: the backbone's backward is a single autograd node (modules._XceptionFn) that writes all
parameter gradients into one flat fp32 arena (executor.GradSink) and reports each parameter as soon as its
wgrad kernel has been *enqueued*.  Parameters are grouped into ~bucket_mb buckets in reverse registration
order (exit flow first = the order backward produces them); when the last parameter of a bucket is reported,
an event is recorded on the compute stream and the bucket's slice of the arena is all-reduced on the
communication stream.  The remaining (LSTM / head) gradients are reduced as one final bucket.  `finish()`
makes the compute stream wait for the communication stream before the optimizer runs.

The in-place reduction of the arena is only the parameters' gradient if autograd *stole* the arena views as ``p.grad``
(``zero_grad(set_to_none=True)`` before every backward).  Two guards keep replicas from silently diverging otherwise:
  * if any backbone ``p.grad`` already exists when backward starts (gradient accumulation as in train_au_face.py:676-693,
    ``set_to_none=False``, retained grads), AccumulateGrad will ADD the arena views into the old tensors on the compute
    stream, so the hooks launch nothing (an in-place reduction would race that add) and ``finish()`` reduces the
    accumulated ``p.grad`` tensors themselves after backward (no overlap, but correct);
  * otherwise ``finish()`` verifies that every backbone ``p.grad`` is a view of the reduced arena and, for one that is not
    (autograd cloned instead of stealing), copies the averaged arena slice over it.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradBucketer:
    def __init__(self, model: torch.nn.Module, backbone: Optional[torch.nn.Module] = None, bucket_mb: float = 25.0,
                 process_group=None):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        import os
        bucket_mb = float(os.environ.get("XCP_DDP_BUCKET_MB", bucket_mb))      # tuning hook
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self.backbone = backbone
        self._avg_supported = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        self._comm_stream = None
        self._plan_key = None
        self._buckets: List[dict] = []
        self._param_bucket: Dict[int, int] = {}
        self.launched: List[tuple] = []          # (lo, hi) ranges all-reduced during the current backward (tests / logging)
        self.direct_reduced = 0                  # gradients finish() had to reduce outside the arena in the last step
        self._last_sink = None
        self._defer = False                      # this backward accumulates into existing p.grad: reduce in finish() instead
        self._tail = {}                          # pre-allocated flat staging buffers (non-backbone / deferred gradients)
        if backbone is not None:
            backbone.__dict__["_grad_ready_hook"] = self._on_ready

    # ------------------------------------------------------------------ bucket plan over the flat arena
    def _plan(self, sink):
        key = (id(sink.params[0]) if sink.params else 0, sink.total)
        if self._plan_key == key:
            for b in self._buckets:
                b["pending"] = b["count"]
            return
        self._plan_key = key
        self._buckets, self._param_bucket = [], {}
        cur = None
        for p in reversed(sink.params):                      # reverse registration order ~ backward order
            o = sink.offsets[id(p)]
            n = p.numel()
            if cur is None or (cur["hi"] - o) > self.bucket_elems:
                cur = {"lo": o, "hi": o + n, "count": 0, "pending": 0}
                self._buckets.append(cur)
            cur["lo"] = min(cur["lo"], o)
            cur["count"] += 1
            self._param_bucket[o] = len(self._buckets) - 1
        for b in self._buckets:
            b["pending"] = b["count"]

    def _all_reduce(self, t: torch.Tensor):
        if self.world == 1:
            return
        if self._avg_supported:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def _launch(self, flat: torch.Tensor, lo: int, hi: int):
        self.launched.append((lo, hi))
        if self.world == 1:
            return
        if flat.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=flat.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(flat.device))
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                self._all_reduce(flat[lo:hi])
        else:
            self._all_reduce(flat[lo:hi])

    # ------------------------------------------------------------------ hook called from GradSink.done()
    def _on_ready(self, sink, lo: int, hi: int):
        if lo < 0:                                           # flush at the end of the backbone's backward
            if self._defer:
                return
            for b in self._buckets:
                if b["pending"] > 0:
                    b["pending"] = 0
                    self._launch(sink.flat, b["lo"], b["hi"])
            return
        if self._plan_key is None or self._plan_key[1] != sink.total or all(b["pending"] == 0 for b in self._buckets):
            self._plan(sink)
        if sink is not self._last_sink:                      # a new backward: the launch log describes one step
            self._last_sink = sink
            self.launched = []
            self._defer = any(p.grad is not None for p in sink.params)
        if self._defer:
            return
        bi = self._param_bucket.get(lo)
        if bi is None:
            return
        b = self._buckets[bi]
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(sink.flat, b["lo"], b["hi"])

    # ------------------------------------------------------------------ after loss.backward()
    def _tail_views(self, grads: List[torch.Tensor]):
        """Flat staging buffer for a fixed list of gradients, allocated once per list signature."""
        key = tuple((g.numel(), g.dtype, g.device) for g in grads)
        ent = self._tail.get(key)
        if ent is None:
            n = sum(g.numel() for g in grads)
            flat = torch.empty((n,), device=grads[0].device, dtype=grads[0].dtype)
            views, off = [], 0
            for g in grads:
                views.append(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            ent = self._tail[key] = (flat, views)
        return ent

    def _reduce_flat(self, grads: List[torch.Tensor]):
        """One all-reduce over `grads` through the staging buffer: multi-tensor copy in, collective, multi-tensor copy out,
        all on the communication stream (the compute stream only records an event)."""
        flat, views = self._tail_views(grads)
        cuda = flat.is_cuda
        if cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=flat.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(flat.device))
            self._comm_stream.wait_event(ev)
        ctx = torch.cuda.stream(self._comm_stream) if cuda else _Null()
        with ctx:
            torch._foreach_copy_(views, grads)
            self.launched.append((0, flat.numel()))
            self._all_reduce(flat)
            torch._foreach_copy_(grads, views)

    def finish(self, extra_params: Optional[List[torch.nn.Parameter]] = None):
        """All-reduce the gradients that did not go through the backbone arena (LSTM, head, ArcFace ...), verify that the
        backbone gradients autograd installed ARE the averaged arena, then join the communication stream."""
        params = extra_params
        bb_params = list(self.backbone.parameters()) if self.backbone is not None else []
        if params is None:
            bb = {id(p) for p in bb_params}
            params = [p for p in self.model.parameters() if id(p) not in bb]
        grads = [p.grad for p in params if p.grad is not None]
        self.direct_reduced = 0
        sink = self._last_sink
        stray = []
        if self.world > 1:
            if self._defer:                                  # accumulated gradients: reduce p.grad itself, after backward
                acc = [p.grad for p in bb_params if p.grad is not None]
                self.direct_reduced = len(acc)
                if acc:
                    self._reduce_flat(acc)
            elif sink is not None:
                lo, hi = sink.flat.data_ptr(), sink.flat.data_ptr() + 4 * sink.flat.numel()
                stray = [p for p in bb_params if p.grad is not None and id(p) in sink.offsets and not (lo <= p.grad.data_ptr() < hi)]
            if grads:
                self._reduce_flat(grads)
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        if stray:                                            # autograd cloned the arena view: install the averaged slice
            self.direct_reduced = len(stray)
            with torch.no_grad():
                torch._foreach_copy_([p.grad for p in stray], [sink.view(p) for p in stray])
        self._last_sink = None
        self._defer = False


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def shard_clips(global_batch: int, rank: int, world: int) -> range:
    """Rank r of N takes clips r::N of the global batch."""
    return range(rank, global_batch, world)
