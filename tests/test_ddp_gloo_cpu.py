"""world_size-2 gloo test (CPU) of the data-parallel host logic: bucket planning over the flat gradient arena,
ready-order launches, flush, and averaging of the non-arena gradients."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeSink:
    def __init__(self, params):
        self.params = params
        self.offsets, off = {}, 0
        for p in params:
            self.offsets[id(p)] = off
            off += (p.numel() + 3) // 4 * 4
        self.total = off
        self.flat = torch.zeros(off)

    def view(self, p):
        o = self.offsets[id(p)]
        return self.flat[o:o + p.numel()].view(p.shape)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_deepfake_detection_b200.ddp import GradBucketer, shard_clips
    torch.manual_seed(0)
    backbone = torch.nn.Sequential(torch.nn.Linear(64, 300), torch.nn.Linear(300, 200), torch.nn.Linear(200, 10))
    head = torch.nn.Linear(10, 1)
    model = torch.nn.Sequential(backbone, head)
    bk = GradBucketer(model, backbone=backbone, bucket_mb=0.05)     # ~13k floats per bucket -> several buckets
    params = list(backbone.parameters())
    for step in range(2):
        sink = _FakeSink(params)
        hook = backbone.__dict__["_grad_ready_hook"]
        for p in reversed(params):                                  # backward order
            sink.view(p).fill_(float(rank + 1) * (1 + step))
            hook(sink, sink.offsets[id(p)], sink.offsets[id(p)] + p.numel())
        hook(sink, -1, -1)
        for p in head.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        bk.finish()
        expect = (1 + 2) / 2.0 * (1 + step)
        for p in params:
            assert torch.allclose(sink.view(p), torch.full_like(p, expect)), "bucketed average wrong"
        for p in head.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, 1.5))
    covered = sorted(bk.launched[:len(bk._buckets)])
    assert len(bk._buckets) >= 3
    assert covered[0][0] == 0 and covered[-1][1] >= sum(p.numel() for p in params)
    assert len(bk.launched) == len(bk._buckets) + 1                # the launch log describes ONE step (+ the head bucket)
    # guard 1: gradient accumulation -- p.grad exists when backward starts, AccumulateGrad will ADD the arena
    # views into it, so nothing may be reduced in place during backward; finish() averages the accumulated p.grad instead
    for p in params:
        p.grad = torch.full_like(p, float(rank + 1))
    sink = _FakeSink(params)
    for p in reversed(params):
        sink.view(p).fill_(100.0 * (rank + 1))
        hook(sink, sink.offsets[id(p)], sink.offsets[id(p)] + p.numel())
    hook(sink, -1, -1)
    assert bk.launched == [] and all(torch.all(sink.view(p) == 100.0 * (rank + 1)) for p in params)
    for p in params:
        p.grad += sink.view(p)                                      # what AccumulateGrad does
    bk.finish([])
    assert bk.direct_reduced == len(params)
    for p in params:
        assert torch.allclose(p.grad, torch.full_like(p, 1.5 + 150.0))
    # guard 2: p.grad was None but autograd cloned the arena view instead of stealing it (its content may be a torn read of
    # the arena being reduced): finish() installs the averaged arena slice
    for p in params:
        p.grad = None
    sink = _FakeSink(params)
    for p in reversed(params):
        sink.view(p).fill_(float(rank + 1))
        hook(sink, sink.offsets[id(p)], sink.offsets[id(p)] + p.numel())
    hook(sink, -1, -1)
    for p in params:
        p.grad = torch.full_like(p, 777.0)
    bk.finish([])
    assert bk.direct_reduced == len(params)
    for p in params:
        assert torch.allclose(p.grad, torch.full_like(p, 1.5))
    assert list(shard_clips(8, rank, world)) == list(range(rank, 8, 2))
    # what the torchrun routes of train_visual.py / train_audio.py rely on: an explicit extra-parameter list (LSTM + head +
    # a separate ArcFace module, some of them frozen = no gradient) and the replica synchronisation helper
    from multimodal_deepfake_detection_b200.loops import broadcast_module_state, init_data_parallel
    arc = torch.nn.Linear(10, 2)
    frozen = torch.nn.Linear(3, 3)
    extra = list(head.parameters()) + list(arc.parameters()) + list(frozen.parameters())
    for p in list(head.parameters()) + list(arc.parameters()):
        p.grad = torch.full_like(p, float(10 * (rank + 1)))
    bk.finish(extra)
    for p in list(head.parameters()) + list(arc.parameters()):
        assert torch.allclose(p.grad, torch.full_like(p, 15.0))
    assert all(p.grad is None for p in frozen.parameters())
    bn = torch.nn.BatchNorm1d(4)
    with torch.no_grad():
        bn.running_mean.fill_(float(rank + 1)); bn.weight.fill_(float(rank + 5))
    broadcast_module_state([bn], buffers_only=True)
    assert torch.all(bn.running_mean == 1.0) and torch.all(bn.weight == float(rank + 5))       # buffers only
    broadcast_module_state([bn])
    assert torch.all(bn.weight == 5.0)
    os.environ.pop("WORLD_SIZE", None)
    assert init_data_parallel() == (1, 0)                           # plain `python script.py`: nothing is touched
    if rank == 0:
        ret.put(len(bk._buckets))
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) >= 3
