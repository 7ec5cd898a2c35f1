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
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self.backbone = backbone
        self._avg_supported = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        self._comm_stream = None
        self._plan_key = None
        self._buckets: List[dict] = []
        self._param_bucket: Dict[int, int] = {}
        self.launched: List[tuple] = []          # (lo, hi) ranges all-reduced, in launch order (for tests / logging)
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
            for b in self._buckets:
                if b["pending"] > 0:
                    b["pending"] = 0
                    self._launch(sink.flat, b["lo"], b["hi"])
            return
        if self._plan_key is None or self._plan_key[1] != sink.total or all(b["pending"] == 0 for b in self._buckets):
            self._plan(sink)
        bi = self._param_bucket.get(lo)
        if bi is None:
            return
        b = self._buckets[bi]
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(sink.flat, b["lo"], b["hi"])

    # ------------------------------------------------------------------ after loss.backward()
    def finish(self, extra_params: Optional[List[torch.nn.Parameter]] = None):
        """All-reduce the gradients that did not go through the backbone arena (LSTM, head, ArcFace ...), then
        join the communication stream."""
        params = extra_params
        if params is None:
            bb = {id(p) for p in self.backbone.parameters()} if self.backbone is not None else set()
            params = [p for p in self.model.parameters() if id(p) not in bb]
        grads = [p.grad for p in params if p.grad is not None]
        if grads and self.world > 1:
            flat = torch.cat([g.reshape(-1) for g in grads])
            self._launch(flat, 0, flat.numel())
            if flat.is_cuda:
                with torch.cuda.stream(self._comm_stream):
                    off = 0
                    for g in grads:
                        g.copy_(flat[off:off + g.numel()].view_as(g))
                        off += g.numel()
            else:
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)


def shard_clips(global_batch: int, rank: int, world: int) -> range:
    """Rank r of N takes clips r::N of the global batch."""
    return range(rank, global_batch, world)
