"""Does NCCL itself find NVLS on this box?  torchrun --nproc-per-node N nccl_nvls_check.py (NCCL_DEBUG=INFO in the env)."""
import os, time
import torch, torch.distributed as dist
dist.init_process_group('nccl')
r = dist.get_rank(); torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
x = torch.ones(11689512, device='cuda', dtype=torch.bfloat16)
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): dist.all_reduce(x)
b.record(); torch.cuda.synchronize()
if r == 0:
    us = a.elapsed_time(b) * 1e3 / 20
    w = dist.get_world_size()
    print(f'NCCL_ALLREDUCE_BF16_23MB us={us:.1f} busbw={2*(w-1)/w*x.numel()*2/us/1e3:.1f} GB/s', flush=True)
dist.destroy_process_group()
