"""All-reduce timing probe: torchrun --nproc-per-node N tools/nccl_probe.py"""
import os
import torch
import torch.distributed as dist

dist.init_process_group("nccl")
rank = dist.get_rank()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
for mb in (1, 25, 100):
    t = torch.randn(mb * 1024 * 1024 // 4, device="cuda")
    for _ in range(5):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        ms = e0.elapsed_time(e1) / 20
        print("all_reduce %4d MB fp32: %.3f ms  (%.1f GB/s bus)" % (mb, ms, 2 * (dist.get_world_size() - 1) / dist.get_world_size() * mb / 1024 / ms * 1e3), flush=True)
dist.destroy_process_group()
