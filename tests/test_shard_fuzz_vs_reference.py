"""Differential fuzz of the data-shard iterator's host side against the INSTALLED, unmodified reference
(`oracle/_ref/dmlcloud/util/data.py`, loaded by path; `xarray` — absent from the image, used only in annotations — comes
from oracle/shims).  Index lists are integers: every comparison is exact."""
import importlib.util
import sys

import numpy as np
import pytest
import torch

from helpers import REPO

REF_DATA = REPO / 'oracle' / '_ref' / 'dmlcloud' / 'util' / 'data.py'
pytestmark = pytest.mark.skipif(not REF_DATA.exists(), reason='oracle/_ref not built (make -C oracle _ref)')


@pytest.fixture(scope='module')
def ref():
    sys.path.insert(0, str(REPO / 'oracle' / 'shims'))
    try:
        spec = importlib.util.spec_from_file_location('_installed_reference_data', REF_DATA)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(str(REPO / 'oracle' / 'shims'))
        sys.modules.pop('xarray', None)
    return mod


def test_shard_indices_random_cases(ref):
    from dmlcloud_b200.util.data import shard_indices

    rng = np.random.RandomState(5)
    for _ in range(400):
        n, world = int(rng.randint(0, 5000)), int(rng.randint(1, 9))
        rank, seed = int(rng.randint(world)), int(rng.randint(0, 2**31 - 1))
        kw = dict(shuffle=bool(rng.randint(2)), even_shards=bool(rng.randint(2)), seed=seed)
        assert shard_indices(n, rank, world, **kw) == ref.shard_indices(n, rank, world, **kw), (n, rank, world, kw)


def test_chunked_sharding_and_sequences_random_cases(ref):
    from dmlcloud_b200.util.data import chunk_and_shard_indices, shard_sequence

    rng = np.random.RandomState(6)
    for _ in range(200):
        n, world, chunk = int(rng.randint(1, 3000)), int(rng.randint(1, 9)), int(rng.randint(1, 64))
        rank, seed = int(rng.randint(world)), int(rng.randint(0, 2**31 - 1))
        kw = dict(shuffle=bool(rng.randint(2)), even_shards=bool(rng.randint(2)), seed=seed)
        assert (chunk_and_shard_indices(n, chunk, rank, world, **kw) ==
                ref.chunk_and_shard_indices(n, chunk, rank, world, **kw)), (n, chunk, rank, world, kw)
        seq = list(range(1000, 1000 + n))
        assert shard_sequence(seq, rank, world, **kw) == ref.shard_sequence(seq, rank, world, **kw)


def test_sharded_sequence_dataset_epochs(ref):
    from dmlcloud_b200.util.data import ShardedSequenceDataset

    rng = np.random.RandomState(7)
    for _ in range(60):
        n, world = int(rng.randint(1, 500)), int(rng.randint(1, 9))
        rank, seed, epoch = int(rng.randint(world)), int(rng.randint(0, 10**6)), int(rng.randint(0, 20))
        kw = dict(shuffle=bool(rng.randint(2)), even_shards=bool(rng.randint(2)), seed=seed, rank=rank, world_size=world)
        mine, theirs = ShardedSequenceDataset(list(range(n)), **kw), ref.ShardedSequenceDataset(list(range(n)), **kw)
        mine.set_epoch(epoch)
        theirs.set_epoch(epoch)
        assert list(iter(mine)) == list(iter(theirs)), (n, kw, epoch)


def test_interleave_batches_random_cases(ref):
    from dmlcloud_b200.util.data import interleave_batches

    rng = np.random.RandomState(8)
    for _ in range(40):
        k, batch, width = int(rng.randint(1, 6)), int(rng.randint(1, 5)), int(rng.randint(1, 4))
        batch *= k  # the reference requires the batch size to be divisible by num_batches
        n_batches = k * int(rng.randint(1, 4))
        g = torch.Generator().manual_seed(int(rng.randint(10**6)))
        data = [torch.randn(batch, width, generator=g) for _ in range(n_batches)]
        got = list(interleave_batches(iter(data), k))
        want = list(ref.interleave_batches(iter(data), k))
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert torch.equal(a, b)
