"""f-5 measurement: what the statistics exchange of synchronised BatchNorm costs per step, three ways
    plain   torch.nn.BatchNorm2d            no exchange at all (the lower bound; different result: per-rank statistics)
    nccl    torch.nn.SyncBatchNorm          all_gather + all_reduce per layer over NCCL (reference pipeline.py:70-71)
    peer    dmlcloud_b200 PeerSyncBatchNorm one libdmlb LL all-reduce each way per layer
on (a) one BatchNorm2d(256) layer, input [32, 256, 14, 14] and (b) torchvision ResNet-18 (20 BN layers), 64 x 3 x 224 x 224
per rank, bf16 autocast, channels-last; forward + backward, no optimizer.  Eager launches, CUDA events on the launching
stream, after warm-up; every rank reports, the slowest counts.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 profiles/syncbn_bench.py > gpurun_out/syncbn_w2.json
"""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e3 / iters], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import torchvision

    from dmlcloud_b200 import _native as N
    from dmlcloud_b200.gradsync import PeerComm
    from dmlcloud_b200.syncbn import convert

    comm = PeerComm(dev, None, max_message_bytes=1 << 20)
    out = {'world': world, 'unit': 'us per forward+backward (max over ranks)'}

    def variants(make):
        torch.manual_seed(0)
        base = make().to(dev)
        return {'plain': copy.deepcopy(base),
                'nccl': torch.nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(base)),
                'peer': convert(copy.deepcopy(base), comm)}

    # (a) one layer
    g = torch.Generator().manual_seed(rank)
    x = torch.randn(32, 256, 14, 14, generator=g).to(dev).requires_grad_(True)
    layer = {}
    for name, m in variants(lambda: torch.nn.BatchNorm2d(256)).items():
        def step(m=m):
            m.zero_grad(set_to_none=True)
            x.grad = None
            m(x).square().mean().backward()
        layer[name] = round(timed(step, 20, 200), 2)
    out['one_layer_256ch'] = layer

    # (b) ResNet-18
    xb = torch.randn(64, 3, 224, 224, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    yb = torch.randint(0, 1000, (64,), generator=g).to(dev)
    net = {}
    launches = {}
    for name, m in variants(torchvision.models.resnet18).items():
        m = m.to(memory_format=torch.channels_last)

        def step(m=m):
            m.zero_grad(set_to_none=True)
            with torch.autocast('cuda', torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(m(xb), yb)
            loss.backward()
        before = N.launch_count()
        net[name] = round(timed(step, 5, 30), 1)
        launches[name] = (N.launch_count() - before) / 35
    out['resnet18_b64'] = net
    out['resnet18_libdmlb_launches_per_step'] = launches
    out['resnet18_exchange_cost_us'] = {k: round(net[k] - net['plain'], 1) for k in ('nccl', 'peer')}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
