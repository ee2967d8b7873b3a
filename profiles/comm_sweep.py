"""Gradient all-reduce sweep on real NVLink: libdmlb's fused peer kernel vs the NCCL route, per bucket size.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/comm_sweep.py > out.json

For every size (the MNIST bucket, the three ResNet-18 DDP buckets, the whole ResNet-18 gradient set) and wire dtype:
  fused   dmlb_comm_allreduce: scale+cast -> peer exchange -> sum -> write-back in ONE kernel.  R launches are captured
          into a CUDA graph and the replay is timed with CUDA events (device time, no host launch latency), max over ranks.
          Besides the default dispatch (algo 0), messages >= 1 MB are also timed with the one-shot (algo 1), two-shot
          (algo 2) and — when the arenas are bound to an NVSwitch multicast object — NVLS (algo 3) kernels forced.
  nccl    what the NCCL route costs for the same result: K1 pack (libdmlb) -> ncclAllReduce (torch.distributed) -> K2
          unpack (libdmlb), timed with CUDA events around 20 back-to-back iterations, max over ranks.
Reported per entry: microseconds, algorithmic bus bandwidth 2(W-1)/W * wire_bytes / time (GB/s) and its fraction of the
measured 770 GB/s per-direction peer bandwidth (B200_PROFILING.md), plus max |fused - nccl| as a cross-check.
"""
import json
import os
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dmlcloud_b200 import _native as N  # noqa: E402
from dmlcloud_b200.gradsync import WIRES, GradBucketSync  # noqa: E402
from dmlcloud_b200.util import distributed as D  # noqa: E402

NVLINK_GBPS = 770.0
SIZES = [('mnist_cnn', 10_330), ('resnet18_b0', 513_000), ('resnet18_b2', 3_963_456), ('resnet18_b1', 7_213_056),
         ('resnet18_all', 11_689_512)]


def main():
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(saved, 'w')
    D.init_process_group_auto()
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device('cuda', torch.cuda.current_device())
    lib = N.cuda_lib(dev.index)
    side = torch.cuda.Stream(device=dev)
    results = []

    def gather_max(x):
        box = [None] * world
        dist.all_gather_object(box, x)
        return max(box)

    for wire in ('bf16', 'fp32'):
        sync = GradBucketSync(dev, wire=wire, route='peer', max_message_bytes=64 << 20)
        multicast = sync.comm.multicast
        for name, n in SIZES:
            g = torch.Generator(device='cpu').manual_seed(1000 + rank)
            base = torch.randn(n, generator=g).to(dev)
            buf = base.clone()
            wire_bytes = n * (2 if wire == 'bf16' else 4)
            # ---- fused peer kernel, graph-timed: R launches captured back to back, replay timed with CUDA events ----
            R = 10

            def time_algo(algo):
                with torch.cuda.stream(side):
                    st = N.stream_ptr(side)
                    N.check(lib.dmlb_comm_allreduce(sync.comm.handle, buf.data_ptr(), n, WIRES[wire], 1.0, None, algo, None, st))
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    dist.barrier()
                    with torch.cuda.graph(graph, stream=side):
                        for _ in range(R):
                            N.check(lib.dmlb_comm_allreduce(sync.comm.handle, buf.data_ptr(), n, WIRES[wire], 1.0, None,
                                                            algo, None, N.stream_ptr(side)))
                    times = []
                    for _ in range(6):
                        dist.barrier()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        graph.replay()
                        b.record()
                        b.synchronize()
                        times.append(a.elapsed_time(b) * 1e3 / R)
                return gather_max(statistics.median(times[1:]))

            with torch.cuda.stream(side):
                N.check(lib.dmlb_comm_allreduce(sync.comm.handle, buf.data_ptr(), n, WIRES[wire], 1.0 / world, None, 0,
                                                None, N.stream_ptr(side)))
                torch.cuda.synchronize()
                fused_result = buf.clone()
                nvls_result = None
                if sync.comm.multicast:
                    buf.copy_(base)
                    N.check(lib.dmlb_comm_allreduce(sync.comm.handle, buf.data_ptr(), n, WIRES[wire], 1.0 / world, None, 3,
                                                    None, N.stream_ptr(side)))
                    torch.cuda.synchronize()
                    nvls_result = buf.clone()
            fused_us = time_algo(0)
            barrier_us = time_algo(5) if wire_bytes <= (256 << 10) else None  # small: LL (auto) vs the barrier one-shot
            big = wire_bytes >= (1 << 20)
            one_us = time_algo(1) if big else None                          # one-shot forced
            pull_us = time_algo(2) if big else None                         # two-shot (peer loads) forced
            nvls_us = time_algo(3) if big and sync.comm.multicast else None  # in-switch reduction forced
            nvls_rs_us = time_algo(4) if big and sync.comm.multicast else None  # in-switch reduce-scatter + peer-load all-gather
            # ---- NCCL route ----
            buf2 = base.clone()
            stage = torch.empty(n, dtype=torch.bfloat16, device=dev) if wire == 'bf16' else None

            def nccl_once(scale):
                s = N.stream_ptr()
                if wire == 'bf16':
                    N.check(lib.dmlb_bucket_pack_f32_bf16(buf2.data_ptr(), stage.data_ptr(), n, scale, s))
                    dist.all_reduce(stage)
                    N.check(lib.dmlb_bucket_unpack_bf16_f32(stage.data_ptr(), buf2.data_ptr(), n, 1.0, None, s))
                else:
                    N.check(lib.dmlb_bucket_scale_f32(buf2.data_ptr(), n, scale, s))
                    dist.all_reduce(buf2)

            nccl_once(1.0 / world)
            torch.cuda.synchronize()
            nccl_result = buf2.clone()
            for _ in range(3):
                nccl_once(1.0)
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                nccl_once(1.0)
            b.record()
            b.synchronize()
            nccl_us = gather_max(a.elapsed_time(b) * 1e3 / 20)
            diff = float((fused_result - nccl_result).abs().max() / nccl_result.abs().max())
            nvls_diff = None if nvls_result is None else float((nvls_result - nccl_result).abs().max() / nccl_result.abs().max())
            bus = 2 * (world - 1) / world * wire_bytes

            def entry(us):
                gbps = bus / (us * 1e-6) / 1e9
                return {'us': round(us, 2), 'bus_GBps': round(gbps, 1), 'frac_of_770': round(gbps / NVLINK_GBPS, 3)}

            results.append({'bucket': name, 'elements': n, 'wire': wire, 'wire_bytes': wire_bytes,
                            'fused_peer_kernel': entry(fused_us),
                            'oneshot_barrier_forced': entry(barrier_us) if barrier_us else None,
                            'oneshot_forced': entry(one_us) if one_us else None,
                            'twoshot_forced': entry(pull_us) if pull_us else None,
                            'nvls_forced': entry(nvls_us) if nvls_us else None,
                            'nvls_reduce_scatter_plus_pull_forced': entry(nvls_rs_us) if nvls_rs_us else None,
                            'rel_diff_nvls_vs_nccl': nvls_diff,
                            'nccl_route_k1_allreduce_k2': entry(nccl_us),
                            'speedup_vs_nccl_route': round(nccl_us / fused_us, 2), 'rel_diff_fused_vs_nccl': diff})
        sync.close()
    if rank == 0:
        print(json.dumps({'world': world, 'gpu': torch.cuda.get_device_name(dev), 'multicast': bool(multicast),
                          'results': results}, indent=1), file=out, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
