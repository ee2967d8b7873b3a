"""The examples run end to end on the GPU: examples/barebone_mnist.py (BASELINE config 1's plumbing — a raw Stage, no
DDP; the reference runs it on CPU, this package has no CPU path so it runs on cuda:0) and examples/mnist.py
(TrainValStage + DDP + captured step), one epoch each on their synthetic data."""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _load(name):
    spec = importlib.util.spec_from_file_location(f'example_{name}', ROOT / 'examples' / f'{name}.py')
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, [f'{name}.py', '1']
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_barebone_mnist_runs_and_learns():
    from dmlcloud_b200.util.distributed import deinitialize_torch_distributed

    mod = _load('barebone_mnist')
    try:
        p = mod.main()
        t = p.tracker
        assert t.epoch == 2 and not p.models  # raw Stage: nothing registered, hence no DDP and no gradient exchange
        assert not p.grad_syncs
        for name in ('train/loss', 'val/loss', 'train/accuracy', 'val/accuracy'):
            assert len(t[name]) == 1 and torch.isfinite(t[name][0])
        assert t['val/accuracy'][0] > 0.5  # the synthetic digits are learnable: one epoch gets well past chance (0.1)
    finally:
        deinitialize_torch_distributed()


def test_dmlcloud_import_alias():
    import dmlcloud_b200.compat  # noqa: F401
    from dmlcloud.metrics import MetricTracker, Reduction
    from dmlcloud.util.distributed import init_process_group_dummy

    assert MetricTracker.__module__ == 'dmlcloud_b200.metrics' and callable(init_process_group_dummy)
    t = MetricTracker()
    t.register_metric('x', Reduction.MAX, globally=False)
    for v in (1.0, 7.0, 3.0):
        t.track('x', torch.tensor(v, device='cuda'))
    t.next_epoch()
    assert t['x'][0].item() == 7.0
