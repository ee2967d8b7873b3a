"""ORACLE — TEST INFRASTRUCTURE ONLY.  The timed CPU baseline: the UNMODIFIED reference, when it is installed.

`oracle/_ref/` holds a `pip install --target` of /root/reference made by `make -C oracle _ref` in the build container
(git-ignored, travels to the GPU box with the snapshot).  This module runs that package's own TrainValStage + DDP over
gloo on CPU tensors (reference stage.py:290-318, pipeline.py:70-75, metrics.py:121-141) on the MNIST CNN of
examples/mnist.py:27-36 with synthetic MNIST-shaped batches, W ranks on the host cores, and times it the way
BASELINE.md §3 defines: samples/s = steps * 32 * W / wall(train_epoch).  The only non-reference code on the path are the
three dependency shims under oracle/shims (progress_table, omegaconf, xarray: absent from the image, not arithmetic).

Two operating points are measured in the same process:
  stock      the reference as it behaves: metrics cross ranks once per epoch (`tracker.next_epoch()` in `_post_epoch`);
             an "epoch" is one window of K steps, windows are repeated until >= `min_seconds` have been timed and the
             MEDIAN window is reported (plus the wall time of the epoch-closing reduce, reported beside it)
  per_step   the stricter operating point the native arm runs at (metrics reduced every step): epochs of ONE step, wall
             time of train_epoch + _reduce_metrics per step

Used by bench.py only (`--impl reference`, and the N=1 `cpu_baseline` leg).  When oracle/_ref is absent the caller falls
back to oracle/ref_port.py (a restatement, `kind: "port"`).  Never imported by the product.
"""
import json
import os
import statistics
import sys
import tempfile
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = HERE / '_ref'
BATCH = 32


def available():
    return (REF / 'dmlcloud' / 'stage.py').exists()


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


WORKLOADS = {  # name -> (samples per rank per step, input shape, classes)
    'mnist': (32, (1, 28, 28), 10),
    'resnet18': (64, (3, 224, 224), 1000),
}


def _batches(seed, steps, workload='mnist'):
    import torch

    batch, shape, classes = WORKLOADS[workload]
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, *shape, generator=g), torch.randint(0, classes, (batch,), generator=g))
            for _ in range(steps)]


def _model_and_optimizer(workload):
    import torch
    from torch import nn

    torch.manual_seed(0)
    if workload == 'mnist':  # examples/mnist.py:27-39
        model = nn.Sequential(nn.Conv2d(1, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                              nn.Conv2d(16, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
                              nn.Linear(784, 10))
        return model, torch.optim.Adam(model.parameters(), lr=1e-3)
    import torchvision

    model = torchvision.models.resnet18()  # BASELINE config 4: 11,689,512 parameters, SGD
    return model, torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)


def _worker(rank, world, initfile, steps, warmup, max_threads, min_seconds, out_path, workload='mnist'):
    os.environ['CUDA_VISIBLE_DEVICES'] = ''  # the CPU path is what is timed: the reference must not find the box's GPUs
    sys.path.insert(0, str(HERE / 'shims'))
    sys.path.insert(1, str(REF))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    import contextlib
    import io
    import warnings

    import torch
    import torch.distributed as dist
    from torch import nn

    warnings.filterwarnings('ignore')
    torch.set_num_threads(1)
    dist.init_process_group('gloo', init_method=f'file://{initfile}', rank=rank, world_size=world)
    from dmlcloud.pipeline import TrainingPipeline  # the reference's own classes
    from dmlcloud.stage import TrainValStage

    assert Path(sys.modules['dmlcloud'].__file__).resolve().is_relative_to(REF.resolve()), 'not the installed reference'

    plan = {'windows': 0, 'singles': 0}

    class RefStage(TrainValStage):
        """examples/mnist.py:14-58 with synthetic data; run_epoch / _reduce_metrics only add stopwatches."""

        def pre_stage(self):
            model, optimizer = _model_and_optimizer(workload)
            self.pipeline.register_model('cnn', model, verbose=False)  # -> DistributedDataParallel (pipeline.py:74)
            self.pipeline.register_optimizer('opt', optimizer)
            self.loss = nn.CrossEntropyLoss()
            self.data = _batches(100 + rank, max(steps, warmup, 16 if workload == 'mnist' else 1), workload)
            self.pipeline.register_dataset('train', self.data[:warmup], verbose=False)
            self.pipeline.register_dataset('val', [], verbose=False)
            self.walls, self.reduces, self.kinds = [], [], []
            self.threads, self.calibration = 1, {}

        def step(self, batch):
            img, target = batch
            output = self.pipeline.models['cnn'](img)
            loss = self.loss(output, target)
            self.track_reduce('accuracy', (output.argmax(1) == target).float().mean())
            return loss

        def table_columns(self):  # (the default layout wants a val/loss column; this run has no validation batches)
            return [{'name': 'Epoch', 'metric': 'misc/epoch'}, {'name': 'Time/Epoch', 'metric': None},
                    {'name': 'Loss', 'metric': 'train/loss'}]

        def _timed_train(self, n):
            self.pipeline.datasets['train'] = self.data[:n]
            dist.barrier()
            t0 = time.perf_counter()
            self.train_epoch()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)

        def run_epoch(self):
            e = self.current_epoch
            if e == 1:  # warm-up + thread calibration: "all the host threads it can use" is not "as many as exist"
                self._timed_train(warmup)
                cands, t = [], 1
                while t < max_threads:
                    cands.append(t)
                    t *= 2
                cands.append(max_threads)
                if workload != 'mnist':  # seconds per step: large convolutions scale with the cores, no sweep needed
                    cands = [max_threads]
                timings = {}
                for n in cands:
                    torch.set_num_threads(n)
                    if workload == 'mnist':
                        self._timed_train(4)  # grows the intra-op pool
                        timings[n] = statistics.median(self._timed_train(16) for _ in range(3))
                    else:
                        timings[n] = self._timed_train(1)
                    if len(timings) >= 3 and timings[n] > 1.8 * min(timings.values()):
                        break
                self.threads = min(timings, key=timings.get)
                self.calibration = {str(k): round(v, 4) for k, v in timings.items()}
                torch.set_num_threads(self.threads)
                per_window = max(self._timed_train(steps), 1e-4)
                plan['windows'] = int(min(200, max(5, -(-min_seconds // per_window) + 2)))
                plan['singles'] = int(min(300, max(20, (min_seconds / 2) // max(per_window / steps, 1e-5))))
                if workload != 'mnist':  # bounded sample: a ResNet-18 step on CPU cores takes seconds
                    plan['windows'], plan['singles'] = 2, 1
                self.kinds.append('warmup')
                self.walls.append(0.0)
                self.pipeline.stages[0].max_epochs = 1 + plan['windows'] + plan['singles']
            elif e <= 1 + plan['windows']:
                self.kinds.append('stock')
                self.walls.append(self._timed_train(steps))
            else:
                self.kinds.append('single')
                self.walls.append(self._timed_train(1))

        def _reduce_metrics(self):
            dist.barrier()
            t0 = time.perf_counter()
            super()._reduce_metrics()  # -> tracker.next_epoch(): 3 gloo collectives per metric (metrics.py:121-141)
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self.reduces.append(float(t))

    pipeline = TrainingPipeline(name='reference-arm')
    stage = RefStage()
    pipeline.append_stage(stage, max_epochs=2)
    with contextlib.redirect_stdout(io.StringIO()):
        pipeline.run()
    if rank == 0:
        stock = [w for w, k in zip(stage.walls, stage.kinds) if k == 'stock']
        stock_red = [r for r, k in zip(stage.reduces, stage.kinds) if k == 'stock']
        single = [w + r for w, r, k in zip(stage.walls, stage.reduces, stage.kinds) if k == 'single']
        med = statistics.median(stock)
        batch = WORKLOADS[workload][0]
        res = {
            'kind': 'reference', 'workload': workload, 'world': world, 'steps': steps, 'threads_per_rank': stage.threads,
            'calibration': stage.calibration, 'windows': len(stock), 'window_seconds_median': med,
            'window_seconds_min': min(stock), 'window_seconds_max': max(stock), 'seconds': sum(stock),
            'samples_per_s': steps * batch * world / med,
            'epoch_reduce_ms': statistics.median(stock_red) * 1e3, 'n_metrics': len(pipeline.tracker.histories),
            'per_step_reduce_samples_per_s': batch * world / statistics.median(single),
            'per_step_reduce_steps': len(single),
        }
        Path(out_path).write_text(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


def run_baseline(world, steps, warmup, total_threads=None, min_seconds=2.0, workload='mnist'):
    """Launch `world` gloo ranks of the installed reference on this host's cores; returns rank 0's timing dict."""
    import torch.multiprocessing as mp

    if not available():
        raise RuntimeError('oracle/_ref is absent: run `make -C oracle _ref` in the build container')
    cores = total_threads or usable_cores()
    threads = max(1, cores // world)
    tmp = tempfile.mkdtemp(prefix='dmlb_ref_')
    out = os.path.join(tmp, 'result.json')
    if workload != 'mnist':
        steps, warmup = min(steps, 2), 1
    mp.spawn(_worker, args=(world, os.path.join(tmp, 'init'), steps, max(3, warmup) if workload == 'mnist' else 1,
                            min(threads, 64), float(min_seconds), out, workload), nprocs=world, join=True)
    res = json.loads(Path(out).read_text())
    res['cores'] = res['threads_per_rank'] * world
    return res


if __name__ == '__main__':
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    print(json.dumps(run_baseline(w, steps=20, warmup=5), indent=1))
