#!/usr/bin/env python
"""bench.py — samples/sec of the MNIST-CNN data-parallel training step through dmlcloud_b200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 20                       # native arm, one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                          # N ranks, one per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W         # the reference's CPU/gloo path (oracle port)

A "step" = one pass of the hot path over one synthetic MNIST-shaped batch (32 samples per rank): zero_grad, forward
(bf16 autocast), backward — DDP hands every gradient bucket to GradBucketSync.hook (libdmlb K1 -> exchange -> K2) —
Adam, 5 tracked metrics folded into the device slab, and the cross-rank metric exchange (fused slab kernel) EVERY step.
Everything goes through the public API: TrainingPipeline.run() -> TrainValStage.train_epoch().

  value   inputs already resident in HBM (K distinct batches), device-timed with CUDA events, max over ranks
  e2e     the same loop fed from pinned HOST memory: H2D copy of every batch and a D2H read of the step's reduced
          metrics (the live exchange copies its result to pinned memory every step; the host reads it one step late)
  roofline            libdmlb bucket kernel (dmlb_bucket_pack_f32_bf16) on a 1 GiB cold buffer, same C-ABI entry point
  roofline_in_situ    the bucket launches inside the timed region (41 KB MNIST bucket: launch-latency bound)
  cpu_baseline        oracle/ref_port.py — the reference's CPU path — on this box's host cores (rank 0, N=1 only)

Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = 'samples/sec (box, device-timed) MNIST CNN'
BATCH = 32
SAMPLE_BYTES_IN = 1 * 28 * 28 * 4  # fp32 image
LABEL_BYTES = 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=30)
    ap.add_argument('--impl', choices=['native', 'reference'], default='native')
    ap.add_argument('--grad-wire', choices=['bf16', 'fp32'], default='bf16')
    ap.add_argument('--grad-route', choices=['auto', 'peer', 'nccl'], default='auto')
    ap.add_argument('--metric-route', choices=['auto', 'peer', 'collective'], default='auto')
    ap.add_argument('--no-graph', action='store_true', help='eager step loop instead of the whole-step CUDA graph')
    ap.add_argument('--adam', choices=['flat', 'torch-fused', 'torch-foreach'], default='flat',
                    help='optimizer of the step: dmlcloud_b200.optim.FlatAdam (libdmlb K5, one launch over flat buffers; '
                         'default, +7%% over torch-fused in the graph step, profiles/README.md) or torch.optim.Adam')
    ap.add_argument('--no-fused-adam', action='store_true', help='same as --adam torch-foreach')
    ap.add_argument('--flat-adam', action='store_true', help='same as --adam flat (the default)')
    ap.add_argument('--channels-last', action='store_true', help='keep model + images in NHWC (cuDNN bf16 native layout)')
    ap.add_argument('--no-micro', action='store_true', help='skip the kernel / metric microbenchmarks')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-steps', type=int, default=3000)
    args = ap.parse_args()
    if args.no_fused_adam:
        args.adam = 'torch-foreach'
    if args.flat_adam:
        args.adam = 'flat'
    return args


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md).  The region can be as short as
    0.1 s, so NVML is polled from a thread every 2 ms (an `nvidia-smi -lms` subprocess would not deliver a single sample
    in that time); nvidia-smi is only the fallback when the NVML bindings are missing."""
    REASONS = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            ids = [v.strip() for v in vis.split(',') if v.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._nvml = None
            return
        self._poll()  # at least one sample even for a very short region
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _poll(self):
        nv = self._nvml
        try:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)))
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle)) if hasattr(
                nv, 'nvmlDeviceGetCurrentClocksEventReasons') else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
            for name, bit in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:  # noqa: BLE001
            pass

    def _run(self):
        while not self._stop.wait(0.002):
            self._poll()

    def stop(self):
        if self._nvml is None:
            return self._smi_once()
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        self._poll()
        return {'sm_mhz': statistics.median(self.samples) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'samples': len(self.samples), 'reasons': sorted(self.reasons), 'how': 'NVML polled every 2 ms'}

    def _smi_once(self):
        try:
            out = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.max.sm', '--format=csv,noheader,nounits',
                                  '-i', str(self._physical_index())], capture_output=True, text=True, timeout=10).stdout
            sm, mx = [float(v) for v in out.strip().split(',')[:2]]
            return {'sm_mhz': sm, 'sm_max_mhz': mx, 'samples': 1, 'reasons': [], 'how': 'nvidia-smi, once, after the region'}
        except Exception:  # noqa: BLE001
            return {'sm_mhz': None, 'sm_max_mhz': None, 'samples': 0, 'reasons': ['no NVML / nvidia-smi']}


# ----------------------------------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------------------------------
def native_arm(args):
    import torch
    import torch.distributed as dist
    from torch import nn

    from dmlcloud_b200 import TrainValStage, _native as N
    from dmlcloud_b200.metrics import Reduction
    from dmlcloud_b200.pipeline import TrainingPipeline
    from dmlcloud_b200.util import distributed as D

    if not torch.cuda.is_available():
        raise SystemExit('bench.py (native arm) needs CUDA: dmlcloud_b200 has no CPU fallback')
    D.init_process_group_auto()
    world, rank = dist.get_world_size(), dist.get_rank()
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE is {world}: launch with torchrun --nproc-per-node {args.gpus}')
    use_graph = not args.no_graph
    # graph mode spends 3 eager steps + 1 capture step before the first replay: keep all of that inside the warm-up
    K, W = args.steps, max(8 if use_graph else 3, args.warmup)
    torch.backends.cudnn.benchmark = True

    def gen_batches(seed, count, pinned):
        g = torch.Generator().manual_seed(seed)
        out = []
        for _ in range(count):
            x = torch.randn(BATCH, 1, 28, 28, generator=g)
            y = torch.randint(0, 10, (BATCH,), generator=g)
            out.append((x.pin_memory(), y.pin_memory()) if pinned else (x, y))
        return out

    class Phase:
        def __init__(self, name, data, timed):
            self.name, self.data, self.timed = name, data, timed
            self.elapsed_ms = None
            self.launches = 0
            self.clocks = None

    class BenchStage(TrainValStage):
        def pre_stage(self):
            dev = self.device
            torch.manual_seed(0)
            model = nn.Sequential(nn.Conv2d(1, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                                  nn.Conv2d(16, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
                                  nn.Linear(784, 10))  # reference examples/mnist.py:27-36
            if args.channels_last:
                model = model.to(memory_format=torch.channels_last)
            self.pipeline.register_model('cnn', model, verbose=False, grad_wire=args.grad_wire)
            if args.adam == 'flat':  # libdmlb K5: parameters, moments and (in graph mode) gradients in flat buffers
                from dmlcloud_b200.optim import FlatAdam

                optimizer = FlatAdam(model.parameters(), lr=1e-3)
            else:
                optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=use_graph,
                                             fused=args.adam == 'torch-fused')
            self.pipeline.register_optimizer('adam', optimizer)
            self.loss = nn.CrossEntropyLoss()
            # whole-step CUDA graph after 3 eager steps (graphstep.py); at W > 1 it needs the peer communicator
            self.cuda_graph = use_graph and (world == 1 or self.pipeline.grad_syncs['cnn'].comm is not None)
            self.live_metrics_every = 1  # metrics cross ranks EVERY step (BASELINE configs 2/3)
            self.tracker.deferred = True
            host = gen_batches(100 + rank, W + K, pinned=True)
            resident = [(x.to(dev), y.to(dev)) for x, y in host]
            self.phases = [Phase('warmup', resident[:W], False), Phase('value', resident[W:], True),
                           Phase('warmup_e2e', host[:W], False), Phase('e2e', host[W:], True)]
            self.pipeline.datasets['train'] = []
            self.pipeline.datasets['val'] = []
            self.host_reads = 0
            self.read_host = False

        def step(self, batch):
            x, y = batch
            x = x.to(self.device, non_blocking=True)  # no-op for the resident phases
            if args.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)  # C == 1: a stride relabel, no copy
            y = y.to(self.device, non_blocking=True)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                out = self.pipeline.models['cnn'](x)
            loss = self.loss(out.float(), y)
            self.track_reduce('accuracy', (out.argmax(1) == y).float().mean())
            return loss

        def table_columns(self):
            return [{'name': 'Epoch', 'metric': 'misc/epoch'}, {'name': 'Time/Epoch', 'metric': None},
                    {'name': 'Loss', 'metric': 'train/loss'}]

        def feed(self, data):
            """The 'DataLoader': hands out the next batch; in the e2e phase it first reads the previous step's
            reduced metrics on the host (the D2H result of the per-step exchange), like a progress bar would."""
            for batch in data:
                if self.read_host and self.live_metrics:
                    self.last_loss = self.live_metrics['train/loss'].value()
                    self.host_reads += 1
                yield batch

        def run_epoch(self):
            phase = self.phases[self.current_epoch - 1]
            self.pipeline.datasets['train'] = self.feed(phase.data)
            self.read_host = phase.name == 'e2e'
            sync = self.pipeline.grad_syncs['cnn']
            sync.profile_events = phase.name == 'value'
            if phase.timed:
                sampler = ClockSampler(self.device.index)
                dist.barrier()
                torch.cuda.synchronize()
                sampler.start()
                n0 = N.launch_count()
                replays0 = self._graph.replays if self._graph is not None else 0
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                wall0 = time.perf_counter()
                t0.record()
            self.train_epoch()  # <- the public per-step loop (stage.py train_epoch), exactly len(phase.data) steps
            if phase.timed:
                t1.record()
                torch.cuda.synchronize()
                wall = (time.perf_counter() - wall0) * 1e3
                dist.barrier()
                phase.elapsed_ms = max(t0.elapsed_time(t1), 0.0)
                phase.wall_ms = wall
                phase.launches = N.launch_count() - n0  # launched through the C ABI in the region ...
                if self._graph is not None:                # ... plus the libdmlb kernels each graph replay re-runs
                    phase.launches += (self._graph.replays - replays0) * self._graph.kernels_in_graph
                phase.clocks = sampler.stop()
            sync.profile_events = False

    pipeline = TrainingPipeline(name='bench')
    pipeline.grad_route, pipeline.metric_route = args.grad_route, args.metric_route
    stage = BenchStage()
    pipeline.append_stage(stage, max_epochs=4)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        pipeline.run()
    dev = pipeline.device
    value_phase, e2e_phase = stage.phases[1], stage.phases[3]

    def max_over_ranks(ms):
        box = [None] * world
        dist.all_gather_object(box, ms)
        return max(box)

    value_ms = max_over_ranks(value_phase.elapsed_ms)
    e2e_ms = max_over_ranks(max(e2e_phase.elapsed_ms, e2e_phase.wall_ms))  # host reads are part of e2e: wall >= device
    samples = K * BATCH * world
    sync = pipeline.grad_syncs['cnn']

    # ---- in-situ bucket kernel timing (events recorded on the launching stream inside the value phase) ----
    torch.cuda.synchronize()
    durs = [a.elapsed_time(b) * 1e3 for a, b, _, _ in sync.event_log]  # us
    n_elem = sync.event_log[0][2] if sync.event_log else 0
    route = sync.event_log[0][3] if sync.event_log else None
    if stage._graph is not None:  # inside the graph no event can be recorded: time the same launches right after
        durs = stage._graph.time_gradient_sync()[3:]
        n_elem = stage._graph.bucket.total
        route = 'single' if world == 1 else 'peer'
    per_elem = {('single', 'bf16'): 8, ('single', 'fp32'): 8, ('peer', 'bf16'): 12, ('peer', 'fp32'): 16,
                ('nccl', 'bf16'): 12, ('nccl', 'fp32'): 8}.get((route, args.grad_wire), 12)
    peaks = load_peaks()
    in_situ = None
    if durs:
        mean_us = statistics.mean(durs)
        achieved = n_elem * per_elem / (mean_us * 1e-6) / 1e9
        in_situ = {'kernel': f'GradBucketSync[{route},{args.grad_wire}] bucket launches', 'elements': n_elem,
                   'bound': 'hbm', 'achieved': round(achieved, 2), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                   'frac': round(achieved / peaks['hbm_gbs'], 5), 'mean_us': round(mean_us, 2),
                   'algorithmic_bytes_per_launch': n_elem * per_elem, 'launches_timed': len(durs),
                   'note': 'MNIST bucket = 10,330 fp32 (41 KB): launch-latency bound, not bandwidth bound'}

    result = {
        'metric': METRIC, 'value': round(samples / (value_ms * 1e-3), 1), 'unit': 'samples/s', 'n_gpus': world,
        'steps': K, 'warmup': W, 'ms_per_step': round(value_ms / K, 4), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16' if args.grad_wire == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': 'MNIST CNN (examples/mnist.py:27-36) DDP, bf16 autocast, Adam, 32 samples/rank/step, '
                               '5 metrics tracked + cross-rank metric exchange every step',
                   'global_batch': BATCH * world, 'parallelism': f'dp{world}', 'grad_wire': args.grad_wire,
                   'cuda_graph': bool(stage._graph is not None), 'channels_last': bool(args.channels_last), 'adam': {'flat': 'libdmlb FlatAdam (K5)', 'torch-fused': 'torch fused', 'torch-foreach': 'torch foreach'}[args.adam],
                   'graph_replays': stage._graph.replays if stage._graph is not None else 0,
                   'grad_route': sorted(set(sync.last_routes.values())),
                   'metric_route': 'peer' if pipeline.metric_comm is not None else ('single' if world == 1 else 'collective'),
                   'l2': 'K distinct batches; the whole working set (<10 MB) is L2-resident by the nature of this '
                         'workload; roofline microbench uses 1 GiB buffers (> 126 MB L2)'},
        'clocks': value_phase.clocks,
        'e2e': {'value': round(samples / (e2e_ms * 1e-3), 1), 'unit': 'samples/s',
                'h2d_bytes_per_step': BATCH * (SAMPLE_BYTES_IN + LABEL_BYTES),
                'd2h_bytes_per_step': 128 + 9 * pipeline.tracker._slab.capacity,
                'ms_per_step': round(e2e_ms / K, 4), 'host_reads': stage.host_reads},
        'gpu_launches': value_phase.launches,
        'wall_ms_per_step': round(value_phase.wall_ms / K, 4),
        'roofline_in_situ': in_situ,
    }

    if rank == 0 and not args.no_micro:
        result.update(kernel_microbench(dev, peaks))
    if not args.no_micro:
        mr = metric_reduce_microbench(pipeline, dev, world, rank)
        if rank == 0:
            result['metric_reduce_us'] = mr
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result['cpu_baseline'] = cpu_baseline(args.cpu_steps)
    dist.barrier()
    if rank == 0:
        print(json.dumps(result), file=JSON_OUT, flush=True)
    for s in pipeline.grad_syncs.values():
        s.close()
    if pipeline.metric_comm is not None:
        pipeline.metric_comm.close()
    dist.destroy_process_group()


def load_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return {'hbm_gbs': float(d['hbm_gbs']), 'source': 'MEASURED_PEAKS.json (of measured)'}
    return {'hbm_gbs': 6650.0, 'source': 'B200_PROFILING.md fallback (of fallback)'}


def ncu_traffic(kernel_substr):
    """DRAM bytes (read + write) per launch of a kernel from the committed `ncu --set full` summary of the same 1 GiB
    launch (profiles/r1_bucket_kernels_ncu_full_final.csv, made by profiles/summarize_ncu.py); None if absent."""
    import csv

    path = ROOT / 'profiles' / 'r1_bucket_kernels_ncu_full_final.csv'
    if not path.exists():
        return None
    rows = list(csv.reader(open(path)))
    head = rows[0]
    try:
        k, t = head.index('Kernel Name'), head.index('dram_bytes_total')
    except ValueError:
        return None
    for r in rows[2:]:
        if kernel_substr in r[k]:
            return int(float(r[t]))
    return None


def kernel_microbench(dev, peaks):
    """The bucket kernels through the C ABI on buffers far larger than L2 (1 GiB fp32 source), CUDA-event timed per launch
    on the launching stream; plus the ResNet-18 bucket sizes with an L2 flush between launches."""
    import torch

    from dmlcloud_b200 import _native as N

    lib = N.cuda_lib(dev.index)
    n = 1 << 28  # 268,435,456 fp32 = 1 GiB
    src = torch.empty(n, dtype=torch.float32, device=dev).normal_()
    wire = torch.empty(n, dtype=torch.bfloat16, device=dev)
    st = N.stream_ptr()

    def timed(fn, reps=10, warm=3, between=None):
        for _ in range(warm):
            fn()
        out = []
        for _ in range(reps):
            if between:
                between()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            out.append(a.elapsed_time(b) * 1e-3)
        return out

    def entry(name, bytes_per_elem, elems, secs, note=None):
        mean = statistics.mean(secs)
        ach = elems * bytes_per_elem / mean / 1e9
        d = {'kernel': name, 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
             'frac': round(ach / peaks['hbm_gbs'], 4), 'traffic': None, 'elements': elems,
             'algorithmic_bytes_per_launch': elems * bytes_per_elem, 'mean_us': round(mean * 1e6, 2),
             'best_us': round(min(secs) * 1e6, 2), 'peak_source': peaks['source']}
        if note:
            d['note'] = note
        return d

    pack = timed(lambda: N.check(lib.dmlb_bucket_pack_f32_bf16(src.data_ptr(), wire.data_ptr(), n, 0.125, st)))
    pack_tma = timed(lambda: N.check(lib.dmlb_bucket_pack_f32_bf16_tma(src.data_ptr(), wire.data_ptr(), n, 0.125, st)))
    pack_regs = timed(lambda: N.check(lib.dmlb_bucket_pack_f32_bf16_regs(src.data_ptr(), wire.data_ptr(), n, 0.125, st)))
    unpack_tma = timed(lambda: N.check(lib.dmlb_bucket_unpack_bf16_f32_tma(wire.data_ptr(), src.data_ptr(), n, 1.0, st)))
    unpack_regs = timed(lambda: N.check(lib.dmlb_bucket_unpack_bf16_f32_regs(wire.data_ptr(), src.data_ptr(), n, 1.0,
                                                                             None, st)))
    unpack = timed(lambda: N.check(lib.dmlb_bucket_unpack_bf16_f32(wire.data_ptr(), src.data_ptr(), n, 1.0, None, st)))
    scale = timed(lambda: N.check(lib.dmlb_bucket_scale_f32(src.data_ptr(), n, 1.0, st)))
    sq = torch.zeros(1, dtype=torch.float64, device=dev)
    read_only = timed(lambda: N.check(lib.dmlb_bucket_sumsq_f32(src.data_ptr(), n, sq.data_ptr(), st)))
    write_only = timed(lambda: src.zero_())  # cudaMemset-class fill by torch: context for the write-heavy kernels
    q = n // 4  # K5 on four quarter-size arrays carved out of `src` (all zero after the fill above: timing only)
    adam_state = torch.zeros(2, dtype=torch.int64, device=dev)
    quarters = [src.data_ptr() + 4 * q * i for i in range(4)]
    adam = timed(lambda: N.check(lib.dmlb_adam_step_f32(quarters[0], quarters[1], quarters[2], quarters[3], q, 1e-3, 0.9,
                                                        0.999, 1e-8, 0.0, 0, 0, None, 0.0, adam_state.data_ptr(), 1, None,
                                                        st)))
    traffic = {'pack': ncu_traffic('pack_bf16_tma_kernel'), 'pack_regs': ncu_traffic('PackBf16'),
               'unpack_tma': ncu_traffic('unpack_bf16_tma_kernel'), 'unpack_regs': ncu_traffic('UnpackBf16'),
               'scale': ncu_traffic('ScaleInplace')}
    out = {
        'hbm_context': {'read_only_GBps': round(n * 4 / statistics.mean(read_only) / 1e9, 1),
                        'write_only_GBps': round(n * 4 / statistics.mean(write_only) / 1e9, 1),
                        'note': 'read-only = dmlb_bucket_sumsq_f32 (4 B/el); write-only = torch zero_ fill (4 B/el); '
                                'the measured copy peak is a 50/50 read/write mix'},
        'roofline': entry('dmlb_bucket_pack_f32_bf16 (K1, default dispatch)', 6, n, pack,
                          'microbench through the same C-ABI entry point on a 1 GiB fp32 source (cold: > 126 MB L2); '
                          'traffic = dram__bytes_read.sum + dram__bytes_write.sum of the same launch from the committed '
                          'ncu --set full capture (profiles/r1_bucket_kernels_ncu_full_final.csv)'),
        'roofline_more': [entry('dmlb_bucket_pack_f32_bf16_tma (K1 via TMA bulk loads)', 6, n, pack_tma),
                          entry('dmlb_bucket_pack_f32_bf16_regs (K1 via LDG.128 x4 in registers)', 6, n, pack_regs),
                          entry('dmlb_bucket_unpack_bf16_f32 (K2, default dispatch)', 6, n, unpack),
                          entry('dmlb_bucket_unpack_bf16_f32_tma (K2 via TMA bulk load + bulk store)', 6, n, unpack_tma),
                          entry('dmlb_bucket_unpack_bf16_f32_regs (K2 via registers)', 6, n, unpack_regs),
                          entry('dmlb_bucket_scale_f32 (K1, fp32 wire, in place)', 8, n, scale),
                          entry('dmlb_adam_step_f32 (K5, Adam step on a flat bucket: 16 B read + 12 B written per element)',
                                28, q, adam)],
    }
    # ResNet-18 DDP buckets.  A single 10 us launch cannot be timed with an event pair (the pair itself
    # costs microseconds), so R x [L2 flush, kernel] is captured into a CUDA graph and timed against R x [L2 flush].
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)  # 256 MB write > 126 MB L2
    side = torch.cuda.Stream(device=dev)

    def graph_time(body, reps=8, R=8):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            body()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                for _ in range(R):
                    body()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3 / R)
        return ts

    def side_ptr():
        return N.stream_ptr(torch.cuda.current_stream(dev))

    base = statistics.median(graph_time(lambda: flush.zero_()))
    buckets = []
    for elems in (513_000, 7_213_056, 3_963_456, 11_689_512):
        s = src[:elems]
        w = wire[:elems]

        def body():
            flush.zero_()
            N.check(lib.dmlb_bucket_pack_f32_bf16(s.data_ptr(), w.data_ptr(), elems, 0.125, side_ptr()))

        secs = [max(t - base, 1e-9) for t in graph_time(body)]
        buckets.append(entry('dmlb_bucket_pack_f32_bf16 (K1)', 6, elems, secs,
                             'cold: 256 MB L2 flush before each launch; graph-captured, flush time subtracted'))
        for label, fn in (('regs', lib.dmlb_bucket_pack_f32_bf16_regs), ('tma', lib.dmlb_bucket_pack_f32_bf16_tma)):
            def body_v(fn=fn):
                flush.zero_()
                N.check(fn(s.data_ptr(), w.data_ptr(), elems, 0.125, side_ptr()))

            secs_v = [max(t - base, 1e-9) for t in graph_time(body_v)]
            buckets.append(entry(f'dmlb_bucket_pack_f32_bf16_{label}', 6, elems, secs_v, f'cold, {label} path forced'))
        # the same bucket as the step sees it: just written by backward, i.e. L2-resident
        warm = graph_time(lambda: N.check(lib.dmlb_bucket_pack_f32_bf16(s.data_ptr(), w.data_ptr(), elems, 0.125,
                                                                         side_ptr())))
        buckets.append(entry('dmlb_bucket_pack_f32_bf16 (K1)', 6, elems, warm,
                             'warm: source L2-resident (as right after backward); back-to-back in a graph'))
    out['roofline_resnet18_buckets'] = buckets
    out['roofline']['traffic'] = traffic['pack']
    for e in out['roofline_more']:
        for key, sub in (('pack_regs', '_regs (K1'), ('pack', '_tma (K1'), ('unpack_tma', 'f32_tma'), ('unpack_regs', 'f32_regs'),
                         ('scale', 'scale_f32')):
            if sub in e['kernel']:
                e['traffic'] = traffic[key]
    del src, wire, flush
    torch.cuda.empty_cache()
    return out


def metric_reduce_microbench(pipeline, dev, world, rank, n_metrics=1024, iters=200, warm=20):
    """BASELINE config 5: 1024 scalar metrics (MEAN/SUM/MIN/MAX mix).  Per step: every metric gets a value, then the
    cross-rank exchange `reduce_live()` runs — ONE fused kernel (finalise + peer exchange + combine) + one D2H copy.
    Reported: CUDA-event time on the launching stream from "last value folded" to "reduced slab copied to pinned host
    memory", and the host wall time of the call; max over ranks.  Also the epoch-closing next_epoch()."""
    import torch
    import torch.distributed as dist

    from dmlcloud_b200.metrics import MetricTracker, Reduction

    ops = [Reduction.MEAN, Reduction.SUM, Reduction.MIN, Reduction.MAX]
    t = MetricTracker()
    t.bind(device=dev, comm=pipeline.metric_comm, group=None)
    t.deferred = True
    names = [f'm{i}' for i in range(n_metrics)]
    for i, name in enumerate(names):
        t.register_metric(name, ops[i % 4])
    vals = torch.randn(n_metrics, generator=torch.Generator().manual_seed(rank)).tolist()
    live_us, live_host_us, epoch_us, pipe_us, dev_us = [], [], [], [], []
    for it in range(warm + iters):
        for name, v in zip(names, vals):
            t.track(name, v)  # python floats ride as kernel immediates (31 per fold launch)
        t._slab.flush()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        a.record()
        live = t.reduce_live()
        b.record()
        h1 = time.perf_counter()
        b.synchronize()
        if it >= warm:
            live_us.append(a.elapsed_time(b) * 1e3)
            live_host_us.append((h1 - h0) * 1e6)
        if it % 10 == 9:
            assert live['m1'].value() is not None
            # steady state: R exchanges back to back (ranks stay coupled through the kernels' own barrier, as in a
            # step loop) -> per-call time without the launch skew a host barrier leaves behind
            R = 10
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(R):
                keep = t.reduce_live()
            p1.record()
            p1.synchronize()
            if it >= warm:
                pipe_us.append(p0.elapsed_time(p1) * 1e3 / R)
            # pure device latency: a peer-barrier kernel first lines the GPUs up in time (every rank has already queued
            # its exchange behind it), so the event pair sees no host launch skew
            if pipeline.metric_comm is not None:
                d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                pipeline.metric_comm.barrier()
                d0.record()
                keep = t.reduce_live()
                d1.record()
                d1.synchronize()
                if it >= warm:
                    dev_us.append(d0.elapsed_time(d1) * 1e3)
            a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a2.record()
            t.next_epoch()
            b2.record()
            b2.synchronize()
            epoch_us.append(a2.elapsed_time(b2) * 1e3)
    t._materialize()

    def stats(xs):
        xs = sorted(xs)
        return {'median': xs[len(xs) // 2], 'p99': xs[max(0, int(len(xs) * 0.99) - 1)], 'min': xs[0]}

    mine = {'live': stats(live_us), 'host': stats(live_host_us), 'epoch': stats(epoch_us), 'pipe': stats(pipe_us),
            'dev': stats(dev_us) if dev_us else stats(pipe_us)}
    box = [None] * world
    dist.all_gather_object(box, mine)

    def worst(kind):
        return {k: round(max(b[kind][k] for b in box), 2) for k in ('median', 'p99', 'min')}

    out = {'n_metrics': n_metrics, 'world': world, 'iters': iters, 'unit': 'us'}
    out.update(worst('live'))
    out['host_call'] = worst('host')
    out['back_to_back'] = worst('pipe')
    out['device_aligned'] = worst('dev')
    out['next_epoch'] = worst('epoch')
    out['what'] = ('median/p99/min: CUDA-event time of MetricTracker.reduce_live() (fused reduce kernel + async D2H of the '
                   'results) with all 1024 metrics holding a value, issued right after a host barrier (includes the ranks\' launch skew); '
                   'back_to_back: per call when 10 exchanges are issued back to back (steady state of a step loop); '
                   'device_aligned (W>1): one exchange queued behind a peer-barrier kernel, i.e. without host launch skew; '
                   'host_call: wall time of the Python call; next_epoch: '
                   'CUDA-event time of the epoch-closing reduce incl. its O(#metrics) host bookkeeping')
    return out


def cpu_baseline(steps):
    """The reference's CPU path (oracle/ref_port.py) on this box's host cores: bounded sample, W=1."""
    from oracle import ref_port

    cores = ref_port.usable_cores()
    res = ref_port.run_baseline(world=1, steps=steps, warmup=50, total_threads=cores, per_step_reduce=True)
    return {'value': round(res['samples_per_s'], 1), 'unit': 'samples/s', 'cores': res['cores'], 'kind': 'port',
            'sample': f'{steps} training steps x 32 samples of the same MNIST-CNN workload (torch CPU, gloo W=1, '
                      f'metrics reduced every step), {res["seconds"]:.1f} s',
            'epoch_reduce_ms': round(res['epoch_reduce_ms'], 3)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm
# ----------------------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle port: torch CPU + DDP/gloo + per-metric gloo
    collectives) on this box's host cores, W = --gpus gloo ranks, all host threads.  Under torchrun only rank 0 works."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_PORT', 'MASTER_ADDR', 'LOCAL_WORLD_SIZE', 'GROUP_RANK'):
        os.environ.pop(k, None)  # the baseline spawns its own gloo world over a file store
    from oracle import ref_port

    world = args.gpus
    cores = ref_port.usable_cores()
    res = ref_port.run_baseline(world=world, steps=args.steps, warmup=max(3, args.warmup), total_threads=cores,
                                per_step_reduce=True)
    value = round(res['samples_per_s'], 1)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': round(res['seconds'] / args.steps * 1e3, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'MNIST CNN (examples/mnist.py:27-36) DDP over gloo on host cores, Adam, 32 samples/rank/'
                               'step, metrics reduced across ranks every step (reference CPU path, oracle/ref_port.py)',
                   'global_batch': BATCH * world, 'parallelism': f'dp{world}', 'threads_per_rank': res['threads_per_rank'],
                   'thread_calibration': res.get('calibration')},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': res['cores'], 'kind': 'port',
                         'sample': f'{args.steps} steps x {BATCH} samples x {world} ranks, {res["seconds"]:.2f} s'},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, cuDNN warnings) write to fd 1 too, so
    fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved original descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, 'w')


if __name__ == '__main__':
    a = parse_args()
    JSON_OUT = _claim_stdout()
    if a.impl == 'reference':
        reference_arm(a)
    else:
        native_arm(a)
