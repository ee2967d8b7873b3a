"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's metric path (/root/reference/dmlcloud/metrics.py):
  Reduction (7-21), reduce_tensor (24-41), MetricReducer (44-155), MetricTracker (158-306).

Differences from the reference are representational only: values are numpy arrays instead of CPU torch tensors, and the
cross-rank step (`dist.all_gather_object` emptiness vote at 124-130, `dist.all_reduce` at 135-140) is expressed over an
explicit list of per-rank reducers (`reduce_across`) so that W>1 behaviour can be evaluated in one process.
dtype rules follow torch.as_tensor: python float -> float32, python int -> int64.

Pinned by tests/test_oracle_pins.py against the reference's own vectors (test/test_metrics.py:9-204) and against
histories produced by the unmodified reference at W=1,2,4 (tests/golden/metrics_w*.json).
"""
from enum import Enum

import numpy as np


class Reduction(Enum):
    MEAN = 'MEAN'
    SUM = 'SUM'
    MIN = 'MIN'
    MAX = 'MAX'


def as_array(value):
    """torch.as_tensor dtype rules on numpy (metrics.py:71)."""
    if isinstance(value, np.ndarray):
        return value
    if isinstance(value, np.generic):  # numpy scalars keep their dtype (np.float64 is also a python float!)
        return np.asarray(value)
    if isinstance(value, bool):
        return np.asarray(value, dtype=np.bool_)
    if isinstance(value, int):
        return np.asarray(value, dtype=np.int64)
    if isinstance(value, float):
        return np.asarray(value, dtype=np.float32)
    if hasattr(value, 'detach'):  # a torch tensor handed in by a test
        return value.detach().cpu().numpy()
    return np.asarray(value)


def reduce_tensor(tensor, reduction, dim=None):
    """metrics.py:24-41."""
    if not isinstance(tensor, np.ndarray):
        raise ValueError('tensor must be an array')
    axes = tuple(range(tensor.ndim)) if dim is None else tuple(dim)
    if reduction is Reduction.MEAN:
        if not np.issubdtype(tensor.dtype, np.floating):
            raise RuntimeError('mean(): input dtype should be floating point')  # what torch raises on int64
        return tensor.mean(axis=axes, dtype=tensor.dtype)
    if reduction is Reduction.SUM:
        out_dtype = np.int64 if tensor.dtype == np.bool_ or np.issubdtype(tensor.dtype, np.integer) else tensor.dtype
        return tensor.sum(axis=axes, dtype=out_dtype)
    if reduction is Reduction.MIN:
        return tensor.min(axis=axes)
    if reduction is Reduction.MAX:
        return tensor.max(axis=axes)
    raise ValueError(f'Unknown reduction {reduction}')


class MetricReducer:
    """metrics.py:44-155."""

    def __init__(self, reduction=Reduction.MEAN, dim=None, globally=True):
        if reduction not in (Reduction.MEAN, Reduction.SUM, Reduction.MIN, Reduction.MAX):
            raise ValueError(f'Unknown reduction {reduction}')
        self.values = []
        self.reduction = reduction
        self.globally = globally
        if isinstance(dim, int):
            self.dim = [dim]
        elif dim is not None:
            self.dim = list(dim)
        else:
            self.dim = None

    def append(self, value):
        self.values.append(as_array(value))

    def clear(self):
        self.values.clear()

    def __len__(self):
        return len(self.values)

    def reduce_locally(self):
        """metrics.py:107-119: stack the step values, reduce over the step axis plus the requested dims."""
        if not self.values:
            return None
        stacked = np.stack(self.values)
        dim = None if self.dim is None else [0] + [d + 1 for d in self.dim]
        return reduce_tensor(stacked, self.reduction, dim=dim)


def reduce_across(reducers):
    """metrics.py:121-141 evaluated for all ranks at once.  reducers: one MetricReducer per rank (rank order).
    Returns the value every rank ends up with (None when empty); raises ValueError on a split emptiness vote."""
    head = reducers[0]
    world = len(reducers)
    if not head.globally:
        raise ValueError('reduce_across is for globally reduced metrics; use reduce_locally per rank otherwise')
    empty = [len(r) == 0 for r in reducers]
    if any(empty):
        if world > 1 and not all(empty):
            raise ValueError('Some workers tracked values this epoch and some did not. This is likely a bug.')
        return None
    local = [r.reduce_locally() for r in reducers]
    if head.reduction is Reduction.MEAN:
        acc = local[0].copy()
        for v in local[1:]:
            acc = (acc + v).astype(acc.dtype)
        return (acc / acc.dtype.type(world)).astype(acc.dtype)  # mean of per-rank means (metrics.py:136-138)
    if head.reduction is Reduction.SUM:
        acc = local[0].copy()
        for v in local[1:]:
            acc = (acc + v).astype(acc.dtype)
        return acc
    if head.reduction is Reduction.MIN:
        return np.minimum.reduce(local)
    return np.maximum.reduce(local)


class MetricTracker:
    """metrics.py:158-306, one instance per simulated rank; `world` (a list of all ranks' trackers, rank order) makes
    reduce_all evaluate the cross-rank step."""

    def __init__(self):
        self.histories = {}
        self.reducers = {}
        self.epoch = 1
        self.world = [self]

    def __contains__(self, name):
        return name in self.histories

    def __len__(self):
        return len(self.histories)

    def __iter__(self):
        return iter(self.histories)

    def _check(self, name):
        if name not in self:
            raise ValueError(f'Metric {name} does not exist')

    def __getitem__(self, name):
        self._check(name)
        return list(self.histories[name])[: self.epoch - 1]

    def has_value(self, name):
        self._check(name)
        return len(self.histories[name]) >= self.epoch

    def current_value(self, name):
        self._check(name)
        return self.histories[name][-1] if self.has_value(name) else None

    def is_reduced_metric(self, name):
        self._check(name)
        return name in self.reducers

    def register_metric(self, name, reduction=None, dim=None, globally=True):
        if name in self:
            raise ValueError(f'Metric {name} already exists')
        if dim is not None and reduction is None:
            raise ValueError('If dim is specified, reduction must be specified as well')
        self.histories[name] = [None] * (self.epoch - 1)
        if reduction is not None:
            self.reducers[name] = MetricReducer(reduction=reduction, dim=dim, globally=globally)

    def track(self, name, value):
        self._check(name)
        if self.has_value(name):
            raise ValueError(f'History for {name} already has a value for epoch {self.epoch}')
        reducer = self.reducers.get(name)
        if reducer is not None:
            reducer.append(value)
        else:
            self.histories[name].append(value)

    def reduce_all(self, prefix=None, strict=True):
        """metrics.py:249-273.  Call on rank 0's tracker when simulating a world: fills every rank's history."""
        for name in list(self.histories):
            if prefix is not None and not name.startswith(prefix):
                continue
            if self.has_value(name):
                if strict:
                    raise ValueError(f'History for {name} has already been reduced for epoch {self.epoch}')
                continue
            reducer = self.reducers.get(name)
            if reducer is None:
                for t in self.world:
                    t.histories[name].append(None)
            elif reducer.globally:
                result = reduce_across([t.reducers[name] for t in self.world])
                for t in self.world:
                    t.histories[name].append(None if result is None else result.copy())
                    t.reducers[name].clear()
            else:
                for t in self.world:
                    t.histories[name].append(t.reducers[name].reduce_locally())
                    t.reducers[name].clear()

    def next_epoch(self):
        self.reduce_all(strict=False)
        for t in self.world:
            t.epoch += 1


def make_world(n):
    trackers = [MetricTracker() for _ in range(n)]
    for t in trackers:
        t.world = trackers
    return trackers
