"""Shared test helpers: replay a golden MetricTracker session, compare histories, spawn gloo ranks."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

from conftest import decode_entry

REPO = Path(__file__).resolve().parent.parent


def replay_metric_script(tracker, script, rank, Reduction, device=None):
    """Run oracle/gen_golden.py's session script on one rank's tracker (product or oracle)."""
    for op in script:
        kind = op[0]
        if kind == 'register':
            tracker.register_metric(op[1], None if op[2] is None else Reduction[op[2]], op[3], op[4])
        elif kind == 'track':
            value = torch.tensor(op[2][rank], dtype=getattr(torch, op[3]))
            tracker.track(op[1], value.to(device) if device is not None else value)
        elif kind == 'track_plain':
            tracker.track(op[1], op[2])
        elif kind == 'reduce_all':
            tracker.reduce_all(prefix=op[1], strict=op[2])
        elif kind == 'next_epoch':
            tracker.next_epoch()


def to_numpy(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return v


def assert_histories_match(histories, epoch, ref_rank, exact_float=False):
    """histories: {name: [tensor|None|py]} from the product; ref_rank: one rank's entry of tests/golden/metrics_w*.json."""
    assert epoch == ref_rank['epoch']
    assert list(histories) == list(ref_rank['histories'])
    for name, ref_hist in ref_rank['histories'].items():
        got_hist = histories[name]
        assert len(got_hist) == len(ref_hist), name
        for want, got in zip(map(decode_entry, ref_hist), got_hist):
            got = to_numpy(got)
            if want is None:
                assert got is None, (name, got)
                continue
            if not isinstance(want, np.ndarray):
                assert got == want, name
                continue
            assert got is not None, name
            assert str(got.dtype) == str(want.dtype), (name, got.dtype, want.dtype)
            assert tuple(got.shape) == tuple(want.shape), (name, got.shape, want.shape)
            exact = np.issubdtype(want.dtype, np.integer) or 'MIN' in name.upper() or 'MAX' in name.upper()
            if exact or exact_float:
                assert (got == want).all(), (name, got, want)  # counters / min / max: bit-exact
            else:
                np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg=name)  # SURVEY §8d tolerance


def spawn(fn, world, *args, timeout=240):
    """Run fn(rank, world, initfile, outdir, *args) in `world` fresh processes; returns the outdir Path."""
    import torch.multiprocessing as mp

    tmp = tempfile.mkdtemp(prefix='dmlb_test_')
    ctx = mp.spawn(fn, args=(world, os.path.join(tmp, 'init'), tmp) + args, nprocs=world, join=False)
    import time

    deadline = time.time() + timeout
    while not ctx.join(timeout=1.0):
        if time.time() > deadline:
            for p in ctx.processes:
                p.kill()
            raise TimeoutError(f'{fn.__name__} did not finish in {timeout}s')
    return Path(tmp)


def init_gloo(rank, world, initfile):
    import torch.distributed as dist

    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / 'tests'))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', init_method=f'file://{initfile}', rank=rank, world_size=world)


def rank_device(rank):
    """CUDA device index for a test rank: distinct GPUs when the box has several (real NVLink peers), cuda:0 shared by
    all ranks on a one-GPU box (peer mappings then go through CUDA IPC on the same device)."""
    n = torch.cuda.device_count()
    return rank % n if n > 0 else 0
