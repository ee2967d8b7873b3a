"""`register_model(sync_bn=True)` (reference pipeline.py:60,70-71 -> torch.nn.SyncBatchNorm) with the statistics
exchange on libdmlb's peer communicator (dmlcloud_b200/syncbn.py).

What SyncBatchNorm computes is, by definition, BatchNorm over the GLOBAL batch: the oracle here is torch's own BatchNorm
(fp32) applied to the concatenation of all ranks' inputs, with autograd — outputs, input gradients, weight / bias gradients
(sum of the ranks' local ones) and running statistics must agree to fp32 round-off.  Ranks are separate processes (real
NVLink peers when the box has several GPUs, CUDA-IPC mappings of one GPU otherwise)."""
import copy
import json
from pathlib import Path

import pytest
import torch

from helpers import init_gloo, rank_device, spawn

pytestmark = pytest.mark.gpu


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                               torch.nn.Conv2d(8, 5, 3, padding=1), torch.nn.BatchNorm2d(5, momentum=0.3))


def _layer_worker(rank, world, initfile, outdir, channels_last):
    init_gloo(rank, world, initfile)
    import torch.distributed as dist

    from dmlcloud_b200 import _native as N
    from dmlcloud_b200.gradsync import PeerComm
    from dmlcloud_b200.syncbn import PeerSyncBatchNorm, convert

    torch.backends.cudnn.allow_tf32 = False
    torch.cuda.set_device(rank_device(rank))
    dev = torch.device('cuda', rank_device(rank))
    comm = PeerComm(dev, None, max_message_bytes=1 << 20)
    ref = _net().to(dev)
    net = convert(copy.deepcopy(ref), comm)
    assert sum(isinstance(m, PeerSyncBatchNorm) for m in net.modules()) == 2
    assert set(net.state_dict()) == set(ref.state_dict())  # same keys: checkpoints are interchangeable
    worst, detail = 0.0, {}
    launches = N.launch_count()
    for step in range(3):
        g = torch.Generator().manual_seed(100 * step + rank)
        x = (torch.randn(4, 3, 8, 8, generator=g) * (1 + rank) + 0.5 * rank).to(dev)  # ranks see different distributions
        w = torch.randn(4, 5, 8, 8, generator=g).to(dev)
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        net.zero_grad()
        out = net(x)
        (out * w).sum().backward()
        # ---- oracle: plain BatchNorm over the global batch ----
        xs, ws = [torch.empty_like(x.detach().contiguous()) for _ in range(world)], [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(xs, x.detach().contiguous())
        dist.all_gather(ws, w)
        gx = torch.cat(xs).requires_grad_(True)
        ref.zero_grad()
        rout = ref(gx)
        (rout * torch.cat(ws)).sum().backward()
        lo, hi = 4 * rank, 4 * rank + 4

        def rel(a, b):
            return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-12))

        detail[f'out/{step}'], detail[f'dx/{step}'] = rel(out, rout[lo:hi]), rel(x.grad, gx.grad[lo:hi])
        # parameter gradients are compared on the scale of the largest one: the bias of a convolution that feeds a BatchNorm
        # has an exactly-zero gradient in exact arithmetic (BN removes the mean), so both sides hold only round-off there
        gscale = max(float(q.grad.abs().max()) for q in ref.parameters())
        for (name, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            total = p.grad.clone()
            dist.all_reduce(total)  # the reference's full-batch parameter gradient is the sum of the ranks' local ones
            detail[f'grad:{name}/{step}'] = float((total - q.grad).abs().max()) / gscale
        for (name, b), (_, c) in zip(net.named_buffers(), ref.named_buffers()):
            if b.dtype.is_floating_point:
                detail[f'buf:{name}/{step}'] = rel(b, c)
            else:
                assert torch.equal(b, c), name  # num_batches_tracked
        worst = max(detail.values())
    per_step = (N.launch_count() - launches) / 3
    net.eval()  # evaluation mode: running statistics, no exchange
    ref.eval()
    before = N.launch_count()
    with torch.no_grad():
        e1, e2 = net(x.detach()), ref(x.detach())
    assert N.launch_count() == before  # no exchange in evaluation mode
    assert rel(e1, e2) < 1e-4, rel(e1, e2)  # (running statistics agree to 2e-5 relative, checked above)
    net.train()
    with pytest.raises(ValueError, match='empty per-rank batch'):  # refused on the host, before any collective is issued
        net(torch.empty(0, 3, 8, 8, device=dev))
    assert N.launch_count() == before
    Path(outdir, f'r{rank}.json').write_text(json.dumps({'worst': worst, 'launches_per_step': per_step,
                                                         'offenders': {k: v for k, v in detail.items() if v > 2e-5}}))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,channels_last', [(2, False), (2, True), (4, False)])
def test_peer_sync_batchnorm_equals_batchnorm_over_the_global_batch(world, channels_last):
    out = spawn(_layer_worker, world, channels_last, timeout=600)
    for r in range(world):
        res = json.loads((out / f'r{r}.json').read_text())
        assert res['worst'] < 2e-5, res
        # two BatchNorm layers: one exchange kernel each in the forward and one each in the backward pass — 4 libdmlb launches
        # per step, where torch.nn.SyncBatchNorm issues 4 NCCL collectives
        assert res['launches_per_step'] == 4, res


def _pipeline_worker(rank, world, initfile, outdir, graph):
    init_gloo(rank, world, initfile)
    import torch.distributed as dist

    from dmlcloud_b200 import TrainValStage
    from dmlcloud_b200.optim import FlatSGD
    from dmlcloud_b200.pipeline import TrainingPipeline
    from dmlcloud_b200.syncbn import PeerSyncBatchNorm
    from dmlcloud_b200.util import distributed as D

    D._here = D.Placement('test', rank, world, rank_device(rank), world, 0)
    torch.cuda.set_device(rank_device(rank))

    class S(TrainValStage):
        def pre_stage(self):
            model = torch.nn.Sequential(_net(), torch.nn.Flatten(), torch.nn.Linear(5 * 8 * 8, 10))
            self.pipeline.register_model('bn', model, sync_bn=True, verbose=False)
            self.pipeline.register_optimizer('sgd', FlatSGD(model.parameters(), lr=0.05, momentum=0.9))
            g = torch.Generator().manual_seed(7 + rank)
            data = [(torch.randn(4, 3, 8, 8, generator=g) + rank, torch.randint(0, 10, (4,), generator=g)) for _ in range(10)]
            self.pipeline.register_dataset('train', data, verbose=False)
            self.pipeline.register_dataset('val', data[:2], verbose=False)
            self.cuda_graph, self.live_metrics_every = graph, int(graph)

        def step(self, batch):
            x, y = batch
            return torch.nn.functional.cross_entropy(self.pipeline.models['bn'](x.to(self.device)), y.to(self.device))

    p = TrainingPipeline(name='syncbn')
    stage = S()
    p.append_stage(stage, max_epochs=2)
    p.run()
    model = p.models['bn']
    assert sum(isinstance(m, PeerSyncBatchNorm) for m in model.modules()) == 2 and p.syncbn_comm is not None
    if graph:
        assert stage._graph is not None and stage._graph.replays == 2 * 10 - 3
    state = torch.cat([t.detach().flatten().double() for t in list(model.parameters()) + [b for b in model.buffers()
                                                                                         if b.dtype.is_floating_point]])
    losses = [float(v) for v in p.tracker['train/loss']]
    Path(outdir, f'r{rank}.json').write_text(json.dumps({'digest': float(state.sum()), 'abs': float(state.abs().sum()),
                                                         'losses': losses}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('graph', [False, True])
def test_register_model_sync_bn_runs_on_the_peer_communicator_w2(graph):
    """Through the public API: DDP + PeerSyncBatchNorm + FlatSGD, eager and captured.  Parameters AND running statistics
    end bit-identical on both ranks (every rank normalises with the same global statistics)."""
    out = spawn(_pipeline_worker, 2, graph, timeout=900)
    res = [json.loads((out / f'r{r}.json').read_text()) for r in range(2)]
    assert res[0]['digest'] == res[1]['digest'] and res[0]['abs'] == res[1]['abs']
    assert all(l == l and l < 10 for l in res[0]['losses']) and res[0]['losses'] == res[1]['losses']
