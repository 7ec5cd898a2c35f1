"""CUDA-graph capture of a whole training step and a host->device input prefetcher.

One XceptionLSTMV training step is ~900 kernel launches (74 convs x fwd/dgrad/wgrad + BN + glue); driven from
Python that is 20-30 ms of host time per step -- as long as the GPU work itself at 4-8 clips per GPU.  The
reference's loops (train_visual.py:563-579, train_audio.py:33-46) are eager PyTorch; here the fixed-shape step
(forward, loss, backward, gradient all-reduce, optimizer) is captured ONCE into a CUDA graph and replayed, so the
host cost per step is one cudaGraphLaunch.  This is plumbing around the C-ABI kernels: every node of the graph is
a kernel of libxcp_sm100.so (plus torch's optimizer / NCCL nodes); nothing is traced or compiled.

Rules for a capturable `step_fn(*static_inputs) -> loss`:
  * no host synchronisation inside (no .item(), no .cpu(), no print of device values);
  * the optimizer must be capturable (torch.optim.Adam(..., capturable=True)) or this package's fused optimizer; pass
    FusedAdam instances as ``optimizers=[...]`` so that replays follow LR-scheduler updates (lr lives in device memory);
  * shapes are static: one GraphedTrainStep per (batch, frames, H, W).
"""
from __future__ import annotations

import gc
from typing import Callable, Iterable, Optional, Sequence

import torch

from ._lib import XcpError


def _bump_pack_caches(modules: Iterable[torch.nn.Module]):
    # Replays update parameters without touching Tensor._version; eager code that runs later (evaluation, checkpoint
    # consumers) must rebuild its bf16 weight packs, so invalidate them.
    for root in modules:
        for m in root.modules():
            c = m.__dict__.get("_pack_cache")
            if c is not None:
                c.generation += 1


class GraphedTrainStep:
    """Capture `step_fn(*inputs)` (a full forward/backward/optimizer step) into a CUDA graph.

    >>> step = GraphedTrainStep(train_step, (clips, labels), modules=[model])
    >>> loss = step(next_clips, next_labels)       # copies into the static buffers, replays, returns the static loss
    """

    def __init__(self, step_fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor],
                 modules: Sequence[torch.nn.Module] = (), warmup: int = 3, zero_grads: Optional[Callable[[], None]] = None,
                 capture_error_mode: str = "thread_local", optimizers: Sequence[object] = ()):
        if not example_inputs or not all(t.is_cuda for t in example_inputs):
            raise XcpError("GraphedTrainStep: example inputs must be CUDA tensors (no CPU path)")
        self.modules = list(modules)
        # optimizers whose lr / weight_decay live in device memory (optim.FusedAdam): their pinned hyper-parameter buffers are
        # refreshed before every replay, so LR schedulers (train_visual.py:534,627; train_au_face.py:620-623) keep working
        self.optimizers = [o for o in optimizers if hasattr(o, "refresh_hyper")]
        self.static_inputs = [t.clone() for t in example_inputs]
        self._step_fn = step_fn
        dev = self.static_inputs[0].device
        # autograd graphs of earlier eager steps (and their AccumulateGrad nodes, which remember the stream they were
        # created on) must be gone before the side-stream warm-up creates the ones the capture will use
        gc.collect()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):                  # lazy initialisation (TMA encoder, optimizer state, NCCL) happens here
                if zero_grads is not None:
                    zero_grads()
                step_fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if zero_grads is not None:
            zero_grads()
        self.graph = torch.cuda.CUDAGraph()
        # "thread_local": backward's kernels are launched from autograd's worker thread into the capturing stream
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.static_loss = step_fn(*self.static_inputs)
        self.replays = 0

    def load_inputs(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)

    def replay(self) -> torch.Tensor:
        for o in self.optimizers:
            o.refresh_hyper()
        self.graph.replay()
        self.replays += 1
        _bump_pack_caches(self.modules)
        return self.static_loss

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if inputs:
            self.load_inputs(*inputs)
        return self.replay()


class GraphedInference:
    """Capture an eval-mode forward (`fn(*inputs) -> tensor or tuple of tensors`) into a CUDA graph (SURVEY.md §8 row f-3;
    the per-clip loop of test_visual.py:609-624).  With BatchNorm folded from the running statistics nothing in the
    forward depends on the batch, so a single clip is ~190 launches of a few microseconds each: eager Python cannot
    issue them as fast as the GPU retires them, one cudaGraphLaunch can.

    >>> infer = GraphedInference(lambda clips: model(model.extract_features(clips)), (clips,), modules=[model])
    >>> probs = infer(next_clips)                  # static output buffer: clone it to keep it across calls

    The modules must be in eval() mode (train-mode BatchNorm would update running statistics at every replay) and their
    parameters must not change between capture and replay (the graph reads the bf16 weight packs made at capture)."""

    def __init__(self, fn: Callable[..., object], example_inputs: Sequence[torch.Tensor], modules: Sequence[torch.nn.Module] = (),
                 warmup: int = 2):
        if not example_inputs or not all(t.is_cuda for t in example_inputs):
            raise XcpError("GraphedInference: example inputs must be CUDA tensors (no CPU path)")
        for root in modules:
            if any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.training for m in root.modules()):
                raise XcpError("GraphedInference: put the model in eval() mode first (train-mode BatchNorm updates its running "
                               "statistics on every replay)")
        self.static_inputs = [t.clone() for t in example_inputs]
        self._fn = fn
        dev = self.static_inputs[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):                  # weight packs, tensor maps and workspaces are built here
                fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_outputs = fn(*self.static_inputs)
        self.replays = 0

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_outputs


class HostPrefetcher:
    """Double-buffered pinned-host -> device staging on a copy stream (what DataLoader(pin_memory=True) +
    .to(device, non_blocking=True) gives the reference's loops, video_dataloader.py:40-51, train_visual.py:564):
    the H2D copy of batch i+1 overlaps the compute of batch i; `next()` hands out device tensors that are ready on
    the current stream."""

    def __init__(self, device: torch.device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self._slots = [None, None]
        self._events = [None, None]
        self._k = 0

    def submit(self, *host_tensors: torch.Tensor):
        """Start copying a batch (pinned host tensors) into the next staging slot."""
        k = self._k
        self._k ^= 1
        # the slot's previous contents may still be read by the consumer's stream
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            if self._slots[k] is None or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(self._slots[k], host_tensors)):
                self._slots[k] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
            for d, h in zip(self._slots[k], host_tensors):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._events[k] = ev
        return k

    def get(self, k: int):
        """Device tensors of slot k, ordered after their copy on the current stream."""
        torch.cuda.current_stream(self.device).wait_event(self._events[k])
        return self._slots[k]
