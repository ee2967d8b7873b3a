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


def run_product(rank, meta, grad_route='auto', metric_route='auto', live_every=0, graph=False, flat_adam=False):
    from dmlcloud_b200 import TrainValStage
    from dmlcloud_b200.pipeline import TrainingPipeline

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    class MNISTStage(TrainValStage):
        def pre_stage(self):
            self.pipeline.register_dataset('train', batches(100 + rank, meta['train_steps'], meta['batch']), verbose=False)
            self.pipeline.register_dataset('val', batches(200 + rank, meta['val_steps'], meta['batch']), verbose=False)
            model = make_cnn()
            self.pipeline.register_model('cnn', model, verbose=False)
            if flat_adam:  # libdmlb K5 instead of the torch optimizer the reference run used
                from dmlcloud_b200.optim import FlatAdam

                optimizer = FlatAdam(model.parameters(), lr=1e-3)
            else:
                optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=graph)
            self.pipeline.register_optimizer('adam', optimizer)
            self.loss = torch.nn.CrossEntropyLoss()
            self.live_metrics_every = live_every
            self.cuda_graph = graph

        def step(self, batch):
            img, target = batch
            img, target = img.to(self.device), target.to(self.device)
            output = self.pipeline.models['cnn'](img)
            loss = self.loss(output, target)
            self.track_reduce('accuracy', (output.argmax(1) == target).float().mean())
            return loss

    p = TrainingPipeline(name='parity')
    p.grad_route, p.metric_route = grad_route, metric_route
    stage = MNISTStage()
    p.append_stage(stage, max_epochs=meta['epochs'])
    p.run()
    params = torch.cat([q.detach().flatten() for q in p.models['cnn'].parameters()]).double()
    return p, stage, float(params.sum()), float(params.abs().sum())


def compare(p, stage, psum, pabs, ref):
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
                else:
                    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-4, err_msg=name)
    np.testing.assert_allclose(psum, ref['param_sum'], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(pabs, ref['param_abs_sum'], rtol=1e-3)


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


def _train_worker(rank, world, initfile, outdir, grad_route, metric_route, graph=False):
    init_gloo(rank, world, initfile)
    import torch.distributed as dist

    from dmlcloud_b200.util import distributed as D

    # all ranks share cuda:0 (one-GPU CI box): placement says local_rank 0 for everybody
    from helpers import rank_device

    D._here = D.Placement('test', rank, world, rank_device(rank), world, 0)
    torch.cuda.set_device(rank_device(rank))
    gold = load_json(f'train_w{world}.json')
    p, stage, psum, pabs = run_product(rank, gold['meta'], grad_route, metric_route, graph=graph)
    compare(p, stage, psum, pabs, gold['ranks'][rank])
    routes = set(p.grad_syncs['cnn'].last_routes.values())
    Path(outdir, f'ok{rank}.json').write_text(json.dumps({'routes': sorted(routes), 'psum': psum}))
    dist.barrier()
    dist.destroy_process_group()


def test_train_w2_peer_path_matches_reference_run():
    """W=2 as two processes on one GPU: gradients through the fused peer all-reduce, metrics through the fused slab
    exchange — against the reference's 2-rank gloo run."""
    out = spawn(_train_worker, 2, 'peer', 'peer', timeout=900)
    res = [json.loads((out / f'ok{r}.json').read_text()) for r in range(2)]
    assert res[0]['routes'] == ['peer'] and res[1]['routes'] == ['peer']
    assert res[0]['psum'] == res[1]['psum']  # replicas stay bit-identical


def test_train_w2_cuda_graph_peer_path_matches_reference_run():
    """W=2 with the captured step: the fused peer all-reduce runs INSIDE the CUDA graph (device-side sequence counter)."""
    out = spawn(_train_worker, 2, 'peer', 'peer', True, timeout=900)
    res = [json.loads((out / f'ok{r}.json').read_text()) for r in range(2)]
    assert res[0]['psum'] == res[1]['psum']
