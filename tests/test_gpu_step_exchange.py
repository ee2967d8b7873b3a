"""GPU parity of the fused step exchange (csrc/peer_comm.cu: one kernel = gradient all-reduce + the step's metric folds +
the cross-rank exchange of the running metric values) through the C ABI, against the numpy oracles:

  gradients   oracle/grad_oracle.py    (rank-ordered fp32 sum of the bf16 / fp32 wire values) — bit-exact
  metrics     oracle/slab_oracle.py    (fold -> finalise -> rank-ordered combine)             — bit-exact

This is the per-step traffic of the reference's hot loop (stage.py:298-314: backward's bucket all-reduce + 4x
track_reduce) at the per-step operating point of BASELINE configs 2/3.  W = 1 runs in-process; W = 2, 4 as separate
processes sharing the visible GPU(s) through CUDA-IPC peer mappings (real NVLink peers when the box has several GPUs).
Also here: a dead peer must poison the step (NaN gradients, TIMEOUT status, host-visible error word) instead of
producing a plausible partial sum.
"""
import ctypes
import json
import time
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import init_gloo, rank_device, spawn

pytestmark = pytest.mark.gpu

N_GRAD = 10_330  # the MNIST CNN's gradient bucket
STEPS = 5


def _desc(op, is_int, globally, f64=False):
    return op | (int(is_int) << 2) | (int(globally) << 3) | (int(f64) << 4)


def _run_steps(rank, world, dev, wire, n_grad, timeout_rank=None):
    """Drive STEPS fused step exchanges through the C ABI and mirror them in the oracles.  Returns a dict of booleans."""
    import torch.distributed as dist

    from dmlcloud_b200 import _native as N
    from dmlcloud_b200.gradsync import WIRES, PeerComm
    from dmlcloud_b200.metrics import STATUS_BYTES, DeviceSlab, HostFeed, StepRing, _layout_hash
    from oracle import grad_oracle
    from oracle.slab_oracle import MAX, MEAN, MIN, SUM, OracleSlab

    lib = N.cuda_lib(dev.index)
    comm = PeerComm(dev, None, max_message_bytes=4 << 20)
    slab = DeviceSlab(dev)
    ora = OracleSlab(capacity=slab.capacity)
    cells = {}
    layout = [('loss', MEAN, False, True, 1), ('acc', MEAN, False, True, 1), ('total', SUM, True, True, 1),
              ('worker', SUM, True, False, 1), ('time', MEAN, False, True, 1), ('vec', MAX, False, True, 4),
              ('low', MIN, False, True, 1)]
    for name, op, is_int, glob, lanes in layout:
        d = _desc(op, is_int, glob)
        cells[name] = slab.alloc(lanes, d)
        assert ora.alloc(lanes, d) == cells[name]
    slab.flush()
    torch.cuda.synchronize()
    glob = [(cells['loss'], cells['total'] + 1), (cells['time'], cells['low'] + 1)]  # two ranges around the local cell
    loc = [(cells['worker'], cells['worker'] + 1)]
    h = _layout_hash(('step-exchange-test', tuple(glob)))
    ring = StepRing(lib, slab.capacity)
    feed = HostFeed(lib)
    feed.assign({cells['time']: (MEAN, False)})
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    rng = np.random.RandomState(100 + rank)
    ok = {'grad_bit_exact': True, 'metrics_bit_exact': True, 'status_ok': True, 'sumsq_ok': True, 'stamps_ok': True}
    st = N.stream_ptr()
    for t in range(1, STEPS + 1):
        g_local = (rng.randn(n_grad) * 3).astype(np.float32)
        bucket = torch.from_numpy(g_local.copy()).to(dev)
        loss = torch.tensor(float(rng.rand()), device=dev)
        acc = torch.tensor(float(rng.rand()), device=dev, dtype=torch.bfloat16)  # a non-fp32 source dtype
        vec = torch.from_numpy(rng.randn(4, 8).astype(np.float32)).to(dev)
        low = torch.tensor(float(rng.randn()), device=dev, dtype=torch.float64)
        step_ms = float(rng.rand() * 3)
        feed.put(cells['time'], step_ms)
        feed.put(cells['time'], step_ms * 0.5)  # two host scalars for one cell between two steps: combined on the host
        feed.commit(t - 1)
        m = N.StepMetrics()
        m.acc, m.cnt, m.desc = slab.acc.data_ptr(), slab.cnt.data_ptr(), slab.desc.data_ptr()
        m.counter, m.out_ring, m.feed = counter.data_ptr(), ring.device_ptr, feed.device_ptr
        m.layout_hash, m.n_cells, m.capacity = h, slab.n_cells, slab.capacity
        m.ring_slots, m.feed_slots = StepRing.SLOTS, HostFeed.SLOTS
        folds = [N.FoldEntry(loss.data_ptr(), 0, N.F32, cells['loss'], 1, 1, 1, 0),
                 N.FoldEntry(acc.data_ptr(), 0, N.BF16, cells['acc'], 1, 1, 1, 0),
                 N.FoldEntry(None, 1, N.F64, cells['total'], 1, 1, 1, 0),
                 N.FoldEntry(None, 2, N.F64, cells['worker'], 1, 1, 2, 0),  # immediate combining two host scalars
                 N.FoldEntry(None, 0, N.SRC_FEED, cells['time'], 1, feed.cols[cells['time']], 1, 0),
                 N.FoldEntry(vec.data_ptr(), 0, N.F32, cells['vec'], 4, 8, 1, 0),
                 N.FoldEntry(low.data_ptr(), 0, N.F64, cells['low'], 1, 1, 1, 0)]
        m.n_folds = len(folds)
        for i, e in enumerate(folds):
            m.folds[i] = e
        ranges = glob + loc
        m.n_ranges, m.n_global_ranges = len(ranges), len(glob)
        for i, (b, e) in enumerate(ranges):
            m.ranges[i] = N.Range(b, e)
        sumsq.zero_()
        if timeout_rank is not None and rank != timeout_rank and t == STEPS:
            break  # this rank "dies" before the last step: the other one must not get a plausible result
        N.check(lib.dmlb_comm_allreduce(comm.handle, bucket.data_ptr(), n_grad, WIRES[wire], 1.0 / world, sumsq.data_ptr(), 0,
                                        ctypes.byref(m), st), 'step exchange')
        # ---- the oracles ----
        ora._fold(cells['loss'], [np.float32(loss.item())])
        ora._fold(cells['acc'], [acc.float().item()])
        ora._fold(cells['total'], [1])
        ora.acc_i[cells['worker']] += 2
        ora.cnt[cells['worker']] += 2
        ora.acc_f[cells['time']] += step_ms + step_ms * 0.5
        ora.cnt[cells['time']] += 2
        for c in range(4):
            ora._fold(cells['vec'] + c, vec[c].cpu().numpy())
        ora._fold(cells['low'], [low.item()])
        if timeout_rank is not None and t == STEPS:  # the survivor: poisoned outputs, error word raised, no hang
            torch.cuda.synchronize()
            status, vals, flags = ring.read(t) if ring.stamp(t) >= t else (None, None, None)
            return {'nan': bool(torch.isnan(bucket).all()), 'status': status, 'failed': comm.failed(),
                    'sumsq_nan': bool(torch.isnan(sumsq).item())}
        pending = ora.reduce(glob, loc, h, reset=False)  # (gloo all_gather_object inside: collective, like the kernel)
        o_status, o_vals, o_flags = pending.get()
        everyone = [None] * world
        dist.all_gather_object(everyone, g_local) if world > 1 else everyone.__setitem__(0, g_local)
        stacked = np.stack(everyone)
        want = grad_oracle.allreduce_f32(stacked) if wire == 'fp32' else grad_oracle.allreduce_bf16(stacked)
        torch.cuda.synchronize()
        got = bucket.cpu().numpy()
        ok['grad_bit_exact'] &= bool((got == want).all())
        ok['sumsq_ok'] &= abs(sumsq.item() - float(np.sum(got.astype(np.float64) ** 2))) <= 1e-12 * max(1.0, sumsq.item())
        ok['stamps_ok'] &= ring.stamp(t) == t and int(counter.item()) == t
        status, vals, flags = ring.read(t)
        ok['status_ok'] &= status == N.METRIC_OK == o_status
        sel = [c for b, e in glob + loc for c in range(b, e)]
        ok['metrics_bit_exact'] &= all(int(vals[c]) == int(o_vals[c]) and int(flags[c]) == int(o_flags[c]) for c in sel)
    if timeout_rank is not None:
        time.sleep(3.0)  # the "dead" rank keeps its arena mapped while the survivor's kernel still signals into it
    comm.close()
    return ok


@pytest.mark.parametrize('wire', ['bf16', 'fp32'])
def test_step_exchange_w1_matches_oracles(wire):
    from dmlcloud_b200 import _native as N

    dev = torch.device('cuda', 0)
    before = N.launch_count()
    ok = _run_steps(0, 1, dev, wire, N_GRAD)
    assert all(ok.values()), ok
    # ONE libdmlb launch per step carries gradients, folds and the metric results (plus slab set-up launches before)
    assert N.launch_count() - before <= STEPS + 12


def _worker(rank, world, initfile, outdir, wire, n_grad, timeout_rank):
    init_gloo(rank, world, initfile)
    import torch.distributed as dist

    torch.cuda.set_device(rank_device(rank))
    dev = torch.device('cuda', rank_device(rank))
    if timeout_rank is not None:
        import dmlcloud_b200.gradsync as G

        orig = G.PeerComm.__init__

        def short(self, *a, **kw):  # a 0.5 s barrier timeout instead of the 10-minute default
            kw['timeout_seconds'] = 0.5
            orig(self, *a, **kw)

        G.PeerComm.__init__ = short
    res = _run_steps(rank, world, dev, wire, n_grad, timeout_rank)
    Path(outdir, f'r{rank}.json').write_text(json.dumps(res))
    if timeout_rank is None:
        dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,wire,n_grad', [(2, 'bf16', N_GRAD), (2, 'fp32', 4099), (4, 'bf16', N_GRAD),
                                               (4, 'bf16', 600_001)])
def test_step_exchange_multi_rank_matches_oracles(world, wire, n_grad):
    """n_grad = 600,001 at W = 4 takes the two-shot all-reduce: the metric CTA rides along there too."""
    out = spawn(_worker, world, wire, n_grad, None, timeout=600)
    for r in range(world):
        res = json.loads((out / f'r{r}.json').read_text())
        if world > 2 and n_grad > 300_000 and wire == 'bf16':
            res.pop('grad_bit_exact')  # two-shot rounds the sum to bf16 for the all-gather half (tests/test_gpu_gradsync.py)
        assert all(res.values()), (r, res)


def test_dead_peer_poisons_the_step_instead_of_hanging():
    """a barrier timeout used to fall through and write a partial sum.  Now: NaN gradients, TIMEOUT metric
    status, a host-visible error word (polled every step by the stage) — and the 10-minute default is configurable."""
    from dmlcloud_b200 import _native as N

    out = spawn(_worker, 2, 'bf16', N_GRAD, 0, timeout=300)
    res = json.loads((out / 'r0.json').read_text())
    assert res['nan'] and res['failed'] and res['sumsq_nan'], res
    assert res['status'] in (N.METRIC_TIMEOUT, None), res
