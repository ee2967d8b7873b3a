"""End-to-end parity on the GPU: the reference's TrainValStage + DDP(gloo, CPU) MNIST-CNN run (tests/golden/train_w*.json,
produced by the unmodified reference) against the same run through dmlcloud_b200 on CUDA.

Counters, epochs and metric names must match exactly; losses / accuracies within rtol 2e-3 (different conv kernels:
cuDNN fp32 with TF32 disabled vs CPU MKL-DNN); after 12 Adam steps the parameter sums within 1e-3.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import decode_entry, load_json
from helpers import init_gloo, spawn

pytestmark = pytest.mark.gpu

NOT_COMPARABLE = ('misc/step_time_ms', 'misc/epoch_time')  # wall-clock values


def make_cnn():
    from torch import nn

    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(1, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                         nn.Conv2d(16, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(), nn.Linear(784, 10))


def batches(seed, steps, batch):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 1, 28, 28, generator=g), torch.randint(0, 10, (batch,), generator=g))
            for _ in range(steps)]


def run_product(rank, meta, grad_route='auto', metric_route='auto', live_every=0, graph=False, flat_adam=False,
                bench_config=False, variant='plain'):
    """The golden run's script through dmlcloud_b200.  bench_config: exactly what bench.py times — bf16 autocast, bf16
    gradient wire, whole-step CUDA graph, FlatAdam, cross-rank metric exchange every step, deferred tracker.
    variant: 'plain' | 'clip' | 'sched' (oracle/gen_golden.py)."""
    from dmlcloud_b200 import TrainValStage
    from dmlcloud_b200.pipeline import TrainingPipeline

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if bench_config:
        graph, flat_adam, live_every = True, True, 1

    class MNISTStage(TrainValStage):
        def pre_stage(self):
            self.pipeline.register_dataset('train', batches(100 + rank, meta['train_steps'], meta['batch']), verbose=False)
            self.pipeline.register_dataset('val', batches(200 + rank, meta['val_steps'], meta['batch']), verbose=False)
            model = make_cnn()
            self.pipeline.register_model('cnn', model, verbose=False, grad_wire='bf16' if bench_config else None)
            if flat_adam:  # libdmlb K5 instead of the torch optimizer the reference run used
                from dmlcloud_b200.optim import FlatAdam

                optimizer = FlatAdam(model.parameters(), lr=1e-3)
            else:
                optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=graph)
            scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=0.5) if variant == 'sched' else None
            self.pipeline.register_optimizer('adam', optimizer, scheduler)
            self.loss = torch.nn.CrossEntropyLoss()
            self.live_metrics_every = live_every
            self.cuda_graph = graph
            self.tracker.deferred = bench_config
            self.step_time_counts, self.live_at_epoch_end = [], []

        def gradient_clip(self):
            return meta.get('gradient_clip', 0.0) if variant == 'clip' else 0.0

        def step(self, batch):
            img, target = batch
            img, target = img.to(self.device), target.to(self.device)
            if bench_config:
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    output = self.pipeline.models['cnn'](img)
                output = output.float()
            else:
                output = self.pipeline.models['cnn'](img)
            loss = self.loss(output, target)
            self.track_reduce('accuracy', (output.argmax(1) == target).float().mean())
            return loss

        def run_epoch(self):
            self.train_epoch()
            # every train step's host-measured step time must have reached its cell exactly once — also when it travels
            # into the captured step through the host feed ring (one replay late, the last one as an immediate)
            slab = self.tracker._slab
            slab.flush_all()
            m = self.tracker.reducers['misc/step_time_ms']
            self.step_time_counts.append(int(slab.cnt[m.cell].item()))
            if self.live_metrics:
                self.live_at_epoch_end.append({k: self.live_metrics[k].value() for k in
                                               ('train/loss', 'train/accuracy', 'misc/total_train_batches')})
            self.val_epoch()

    p = TrainingPipeline(name='parity')
    p.grad_route, p.metric_route = grad_route, metric_route
    stage = MNISTStage()
    p.append_stage(stage, max_epochs=meta['epochs'])
    p.run()
    params = torch.cat([q.detach().flatten() for q in p.models['cnn'].parameters()]).double()
    assert stage.step_time_counts == [meta['train_steps']] * meta['epochs'], stage.step_time_counts
    return p, stage, float(params.sum()), float(params.abs().sum())


# Tolerances (stated here and in DESIGN.md §3 "Numerics"):
#   fp32 runs       losses / accuracies rtol 2e-3, atol 1e-4; parameter sums 1e-3 — cuDNN fp32 convolutions against the
#                   CPU's MKL-DNN over 12 Adam steps (SURVEY §8d's 1e-5 applies to the metric REDUCTION, which is tested
#                   bit-exactly against the oracle; these runs also contain the user's model)
#   bench config    bf16 autocast forward/backward + bf16 gradient wire against the reference's fp32 CPU run: losses
#                   rtol 2e-2, accuracies atol 4e-2 (an argmax flip moves a 192-sample mean by 5e-3), parameter sums 2e-2
#   always          integer counters, epochs, metric names and history lengths: bit-exact
def compare(p, stage, psum, pabs, ref, loose=False):
    assert p.tracker.epoch == ref['tracker_epoch'] and stage.current_epoch == ref['stage_epoch']
    hist = p.tracker.histories
    assert set(hist) == set(ref['histories'])
    for name, ref_hist in ref['histories'].items():
        assert len(hist[name]) == len(ref_hist), name
        if name in NOT_COMPARABLE:
            continue
        for got, want in zip(hist[name], map(decode_entry, ref_hist)):
            if want is None:
                assert got is None, name
            elif not isinstance(want, np.ndarray):
                assert got == want, name
            else:
                got = got.numpy()
                assert str(got.dtype) == str(want.dtype) and got.shape == want.shape, name
                if np.issubdtype(want.dtype, np.integer):
                    assert (got == want).all(), name  # step / batch counters: bit-exact
                elif loose:
                    np.testing.assert_allclose(got, want, rtol=2e-2, atol=4e-2 if 'accuracy' in name else 1e-3, err_msg=name)
                else:
                    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-4, err_msg=name)
    tol = 2e-2 if loose else 1e-3
    np.testing.assert_allclose(psum, ref['param_sum'], rtol=tol, atol=tol)
    np.testing.assert_allclose(pabs, ref['param_abs_sum'], rtol=tol)


def check_live_equals_history(p, stage):
    """The last step's live exchange (running value of the epoch so far, from the fused step exchange's result ring) covers
    exactly the values the epoch-closing reduce covers: same cells, same finalise, same rank-ordered combine -> every bit."""
    assert len(stage.live_at_epoch_end) == stage.current_epoch - 1
    for epoch, live in enumerate(stage.live_at_epoch_end):
        for name, value in live.items():
            want = p.tracker.histories[name][epoch]
            assert value is not None and torch.equal(value, want), (epoch, name, value, want)


def test_train_w1_matches_reference_run():
    from dmlcloud_b200 import _native as N
    from dmlcloud_b200.util.distributed import deinitialize_torch_distributed, init_process_group_dummy

    gold = load_json('train_w1.json')
    init_process_group_dummy()
    try:
        before = N.launch_count()
        p, stage, psum, pabs = run_product(0, gold['meta'], live_every=2)
        compare(p, stage, psum, pabs, gold['ranks'][0])
        assert N.launch_count() - before > 40  # the libdmlb kernels really ran (bucket + metric launches)
        assert p.grad_syncs['cnn'].buckets_seen >= gold['meta']['train_steps'] * gold['meta']['epochs']
        assert stage.live_metrics and stage.live_metrics['train/loss'].value() is not None
    finally:
        deinitialize_torch_distributed()


def test_train_w1_cuda_graph_step_matches_reference_run():
    """The whole-step CUDA graph (3 eager warm-up steps, capture, replays) reproduces the reference run as well."""
    from dmlcloud_b200.util.distributed import deinitialize_torch_distributed, init_process_group_dummy

    gold = load_json('train_w1.json')
    init_process_group_dummy()
    try:
        p, stage, psum, pabs = run_product(0, gold['meta'], graph=True)
        compare(p, stage, psum, pabs, gold['ranks'][0])
        assert stage._graph is not None and stage._graph.replays == gold['meta']['train_steps'] * gold['meta']['epochs'] - 3
        assert stage._graph.bucket.attached()
    finally:
        deinitialize_torch_distributed()


@pytest.mark.parametrize('graph', [False, True])
def test_train_w1_flat_adam_matches_reference_run(graph):
    """Same golden run with `optimizer.step()` on libdmlb (FlatAdam, K5): per-parameter launches in the eager loop, one
    launch on the flat buffers inside the captured step."""
    from dmlcloud_b200 import _native as N
    from dmlcloud_b200.util.distributed import deinitialize_torch_distributed, init_process_group_dummy

    gold = load_json('train_w1.json')
    steps = gold['meta']['train_steps'] * gold['meta']['epochs']
    init_process_group_dummy()
    try:
        before = N.launch_count()
        p, stage, psum, pabs = run_product(0, gold['meta'], graph=graph, flat_adam=True)
        compare(p, stage, psum, pabs, gold['ranks'][0])
        assert p.optimizers['adam'].steps_taken() == steps
        if graph:
            assert stage._graph is not None and stage._graph.replays == steps - 3 and stage._graph.bucket.attached()
        else:
            assert N.launch_count() - before >= 6 * steps  # six parameter tensors, one K5 launch each per step
    finally:
        deinitialize_torch_distributed()


def _w1(fn):
    from dmlcloud_b200.util.distributed import deinitialize_torch_distributed, init_process_group_dummy

    init_process_group_dummy()
    try:
        return fn()
    finally:
        deinitialize_torch_distributed()


def test_train_w1_bench_configuration_matches_reference_run():
    """the EXACT configuration bench.py times — bf16 autocast, bf16 wire, captured step with the fused
    step exchange (live metrics every step), FlatAdam, deferred tracker — against the reference's run."""
    from dmlcloud_b200 import _native as N

    gold = load_json('train_w1.json')
    steps = gold['meta']['train_steps'] * gold['meta']['epochs']

    def body():
        before = N.launch_count()
        p, stage, psum, pabs = run_product(0, gold['meta'], bench_config=True)
        compare(p, stage, psum, pabs, gold['ranks'][0], loose=True)
        check_live_equals_history(p, stage)
        g = stage._graph
        assert g is not None and g.replays == steps - 3 and g.step_metrics is not None and g.feed is not None
        # the captured step launches THREE libdmlb kernels: the fused step exchange (all-reduce + folds + metric exchange),
        # the K5 optimizer step — and nothing per step outside the graph
        assert g.kernels_in_graph == 2, g.kernels_in_graph
        eager = 3 * (1 + 1 + 6 + 1)  # 3 warm-up steps: bucket launch, fold launch, 6 per-parameter K5, live exchange
        assert N.launch_count() - before < eager + 2 * gold['meta']['epochs'] * 4 + 40
        assert p.optimizers['adam'].steps_taken() == steps

    _w1(body)


@pytest.mark.parametrize('graph', [False, True])
def test_train_w1_gradient_clipping_matches_reference_run(graph):
    """reference stage.py:276-285 with gradient_clip() != 0, through the stage: eager = the all-reduce's fused sum of
    squares + one scale pass (no extra read of the gradients); captured = coefficient applied inside the K5 launch."""
    from dmlcloud_b200 import _native as N

    gold = load_json('train_clip_w1.json')

    def body():
        before = N.launch_count()
        p, stage, psum, pabs = run_product(0, gold['meta'], graph=graph, flat_adam=graph, variant='clip')
        compare(p, stage, psum, pabs, gold['ranks'][0])
        if not graph:
            sync = p.grad_syncs['cnn']
            assert sync.sumsq is not None  # the bucket launches accumulated it; no dmlb_bucket_sumsq pass was needed
            steps = gold['meta']['train_steps'] * gold['meta']['epochs']
            # per step: the bucket launch(es) of the hook (W = 1, fp32 wire: scale + sum of squares of the ONE bucket), 6 clip
            # launches (one per parameter tensor), 1 fold — and no per-parameter sum-of-squares pass (6 more launches)
            assert N.launch_count() - before <= steps * 9 + 40

    _w1(body)


def test_train_w1_scheduler_is_followed_by_the_captured_step():
    """a python-float lr was baked into the captured graph.  FlatAdam keeps lr in device memory: StepLR halves
    it after every epoch and the replays follow (reference stage.py:316-318)."""
    gold = load_json('train_sched_w1.json')

    def body():
        p, stage, psum, pabs = run_product(0, gold['meta'], graph=True, flat_adam=True, variant='sched')
        compare(p, stage, psum, pabs, gold['ranks'][0])
        assert [float(v) for v in p.tracker['misc/lr_adam']] == [1e-3, 5e-4, 2.5e-4]
        opt = p.optimizers['adam']
        assert float(opt._flat[0]['lr'].item()) == 2.5e-4   # device-resident value the last epoch's replays applied
        assert opt.param_groups[0]['lr'] == 1.25e-4           # the scheduler's next value, synced before the next replay

    _w1(body)


def test_torch_optimizer_with_scheduler_is_refused_in_graph_mode():
    gold = load_json('train_sched_w1.json')

    def body():
        with pytest.raises(RuntimeError, match='learning rate is baked'):
            run_product(0, gold['meta'], graph=True, flat_adam=False, variant='sched')

    _w1(body)


def _train_worker(rank, world, initfile, outdir, grad_route, metric_route, graph=False, bench_config=False,
                  variant='plain'):
    init_gloo(rank, world, initfile)
    import torch.distributed as dist

    from dmlcloud_b200.util import distributed as D

    # all ranks share cuda:0 (one-GPU CI box): placement says local_rank 0 for everybody
    from helpers import rank_device

    D._here = D.Placement('test', rank, world, rank_device(rank), world, 0)
    torch.cuda.set_device(rank_device(rank))
    gold = load_json({'plain': f'train_w{world}.json', 'clip': f'train_clip_w{world}.json'}[variant])
    p, stage, psum, pabs = run_product(rank, gold['meta'], grad_route, metric_route, graph=graph,
                                       flat_adam=(graph and variant == 'clip'), bench_config=bench_config, variant=variant)
    compare(p, stage, psum, pabs, gold['ranks'][rank], loose=bench_config)
    if bench_config:
        check_live_equals_history(p, stage)
        assert stage._graph.kernels_in_graph == 2
    routes = set(p.grad_syncs['cnn'].last_routes.values())
    Path(outdir, f'ok{rank}.json').write_text(json.dumps({'routes': sorted(routes), 'psum': psum}))
    dist.barrier()
    dist.destroy_process_group()


def _same_replicas(out, world):
    res = [json.loads((out / f'ok{r}.json').read_text()) for r in range(world)]
    assert all(r['psum'] == res[0]['psum'] for r in res)  # replicas stay bit-identical
    return res


def test_train_w2_peer_path_matches_reference_run():
    """W=2 as two processes on one GPU: gradients through the fused peer all-reduce, metrics through the fused slab
    exchange — against the reference's 2-rank gloo run."""
    res = _same_replicas(spawn(_train_worker, 2, 'peer', 'peer', timeout=900), 2)
    assert res[0]['routes'] == ['peer'] and res[1]['routes'] == ['peer']


def test_train_w2_cuda_graph_peer_path_matches_reference_run():
    """W=2 with the captured step: the fused peer all-reduce runs INSIDE the CUDA graph (device-side sequence counter)."""
    _same_replicas(spawn(_train_worker, 2, 'peer', 'peer', True, timeout=900), 2)


@pytest.mark.parametrize('world', [2, 4, 8])
def test_train_bench_configuration_multi_rank_matches_reference_run(world):
    """The benched configuration at W = 2, 4, 8 (golden runs of the unmodified reference at the same W): ONE peer barrier
    per step carries the gradients and the metric records; the live ring equals the epoch histories bit for bit."""
    _same_replicas(spawn(_train_worker, world, 'peer', 'peer', True, True, timeout=1500), world)


@pytest.mark.parametrize('graph', [False, True])
def test_train_w2_gradient_clipping_matches_reference_run(graph):
    _same_replicas(spawn(_train_worker, 2, 'peer', 'peer', graph, False, 'clip', timeout=900), 2)


def test_staged_host_batches_give_the_same_run_as_resident_batches():
    """graphstep._load: large pinned host batches reach the captured step through a copy stream + two staging buffers
    (the H2D transfer overlaps the previous step).  Same data, resident vs pinned-host: bit-identical parameters."""
    from dmlcloud_b200 import TrainValStage
    from dmlcloud_b200.optim import FlatSGD
    from dmlcloud_b200.pipeline import TrainingPipeline

    def run(pinned):
        g = torch.Generator().manual_seed(5)
        data = [(torch.randn(64, 8192, generator=g), torch.randint(0, 10, (64,), generator=g)) for _ in range(12)]  # 2 MB each
        data = [(x.pin_memory(), y.pin_memory()) if pinned else (x.cuda(), y.cuda()) for x, y in data]

        class S(TrainValStage):
            def pre_stage(self):
                torch.manual_seed(0)
                model = torch.nn.Sequential(torch.nn.Linear(8192, 64), torch.nn.Tanh(), torch.nn.Linear(64, 10))
                self.pipeline.register_model('m', model, verbose=False)
                self.pipeline.register_optimizer('sgd', FlatSGD(model.parameters(), lr=0.05, momentum=0.9))
                self.pipeline.register_dataset('train', data, verbose=False)
                self.pipeline.register_dataset('val', [], verbose=False)
                self.cuda_graph = True
                self.live_metrics_every = 1

            def step(self, batch):
                x, y = batch
                return torch.nn.functional.cross_entropy(self.pipeline.models['m'](x.to(self.device)), y.to(self.device))

            def table_columns(self):
                return [{'name': 'Epoch', 'metric': 'misc/epoch'}, {'name': 'Loss', 'metric': 'train/loss'}]

        p = TrainingPipeline(name='staged')
        stage = S()
        p.append_stage(stage, max_epochs=2)
        p.run()
        assert stage._graph is not None and bool(stage._graph._staging) == pinned
        return torch.cat([q.detach().flatten() for q in p.models['m'].parameters()]).cpu(), p.tracker['train/loss']

    def body():
        a, la = run(False)
        b, lb = run(True)
        assert torch.equal(a, b) and all(torch.equal(x, y) for x, y in zip(la, lb))

    _w1(body)


def test_captured_step_refuses_per_group_clipping_it_cannot_honour():
    """The reference clips per optimizer param group (stage.py:276-279); the captured step's fused sum of squares covers the
    whole flat bucket, so with two groups it must refuse instead of silently clipping differently."""
    from dmlcloud_b200 import TrainValStage
    from dmlcloud_b200.optim import FlatAdam
    from dmlcloud_b200.pipeline import TrainingPipeline

    class S(TrainValStage):
        def pre_stage(self):
            model = make_cnn()
            self.pipeline.register_model('cnn', model, verbose=False)
            head = list(model[-1].parameters())
            body = [p for p in model.parameters() if all(p is not h for h in head)]
            self.pipeline.register_optimizer('adam', FlatAdam([{'params': body}, {'params': head, 'lr': 1e-4}], lr=1e-3))
            self.pipeline.register_dataset('train', batches(1, 6, 8), verbose=False)
            self.pipeline.register_dataset('val', [], verbose=False)
            self.cuda_graph = True

        def gradient_clip(self):
            return 1.0

        def step(self, batch):
            x, y = batch
            return torch.nn.functional.cross_entropy(self.pipeline.models['cnn'](x.to(self.device)), y.to(self.device))

        def table_columns(self):
            return [{'name': 'Epoch', 'metric': 'misc/epoch'}]

    def body():
        p = TrainingPipeline(name='refuse')
        p.append_stage(S(), max_epochs=1)
        with pytest.raises(RuntimeError, match='exactly one optimizer param group'):
            p.run()

    _w1(body)
