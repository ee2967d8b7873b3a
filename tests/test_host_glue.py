"""Host glue around the hot path that needs no GPU: the per-epoch table the stages fill (reference stage.py:147-205 hands user
code the ProgressTable itself), the `import dmlcloud` alias for code written against the reference, and the
snapshot writer (f-2) on a CPU device — the same writer thread, atomic rename and error reporting the GPU path uses, minus
the device -> pinned staging."""
import io
import subprocess
import sys
import threading
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent


class TestEpochTable:
    def _table(self, sink=None):
        from dmlcloud_b200.util.table import EpochTable

        spec = ['train/loss', {'name': 'Epoch', 'metric': 'misc/epoch'}, {'name': 'ETA', 'metric': None, 'width': 9}]
        return EpochTable(spec, sink if sink is not None else io.StringIO())

    def test_reference_style_calls_reach_the_backend(self):
        """stage.py:147,159,168,192,195-205: add_column / __setitem__ / update / next_row / close on `self.table`"""
        sink = io.StringIO()
        t = self._table(sink)
        t['Epoch'] = 3
        assert t['Epoch'] == 3
        t.update('ETA', '0:01:00')
        t.add_column('lr', width=8)
        assert t.has('lr') and t.has('ETA') and not t.has('nope')
        t.add_column('lr')  # twice: still one column
        assert [c['name'] for c in t.columns].count('lr') == 1
        t.update('lr', 0.5)
        t.update('train/loss', 1.25)
        t.next_row()
        t.close()
        header, row = sink.getvalue().strip().splitlines()
        assert [c.strip() for c in header.split('|')] == ['train/loss', 'Epoch', 'ETA', 'lr']
        assert [c.strip() for c in row.split('|')] == ['1.2500', '3', '0:01:00', '0.5000']

    def test_emit_row_takes_the_latest_epoch_of_each_metric_column(self):
        sink = io.StringIO()
        t = self._table(sink)
        history = {'train/loss': [2.0, torch.tensor(1.5)], 'misc/epoch': [1, 2]}
        t.set('ETA', 'soon')
        t.emit_row(history)
        row = sink.getvalue().strip().splitlines()[-1]
        assert [c.strip() for c in row.split('|')] == ['1.5000', '2', 'soon']

    def test_bad_column_specs_raise_like_the_reference(self):
        from dmlcloud_b200.util.table import EpochTable

        with pytest.raises(ValueError, match='Must be a string or a dict'):
            EpochTable([3], io.StringIO())
        with pytest.raises(ValueError, match='"metric" key'):
            EpochTable([{'name': 'x'}], io.StringIO())


class TestDmlcloudAlias:
    def _run(self, code):
        return subprocess.run([sys.executable, '-c', code], cwd=REPO, capture_output=True, text=True, timeout=300)

    def test_reference_imports_resolve_to_this_package(self):
        res = self._run(
            'import dmlcloud_b200.compat\n'
            'import dmlcloud, dmlcloud_b200\n'
            'from dmlcloud import Stage, TrainValStage\n'
            'from dmlcloud.pipeline import TrainingPipeline\n'
            'from dmlcloud.metrics import MetricTracker, MetricReducer, Reduction, reduce_tensor\n'
            'from dmlcloud.checkpoint import CheckpointDir, generate_checkpoint_path, find_slurm_checkpoint, generate_id, '
            'sanitize_filename\n'
            'from dmlcloud.util.distributed import init_process_group_auto, init_process_group_dummy, is_root, root_only, '
            'root_first, rank, world_size, local_rank, all_gather_object, broadcast_object, print_root\n'
            'from dmlcloud.util.data import shard_indices, chunk_and_shard_indices, shard_sequence, ShardedSequenceDataset, '
            'PrefetchDataset, BatchDataset, interleave_batches\n'
            'assert dmlcloud is dmlcloud_b200 and TrainingPipeline is dmlcloud_b200.pipeline.TrainingPipeline\n'
            'import dmlcloud_b200.compat as c; c.install()  # idempotent\n'
            'print("alias ok")\n')
        assert res.returncode == 0 and 'alias ok' in res.stdout, res.stderr[-2000:]

    def test_alias_refuses_to_shadow_another_dmlcloud(self):
        res = self._run(
            'import sys, types\n'
            'sys.modules["dmlcloud"] = types.ModuleType("dmlcloud")  # e.g. the installed reference\n'
            'try:\n'
            '    import dmlcloud_b200.compat\n'
            'except ImportError as exc:\n'
            '    print("refused:", exc)\n')
        assert res.returncode == 0 and 'refused:' in res.stdout, res.stderr[-2000:]


class TestSnapshotWriterOnCpu:
    def _dir(self, tmp_path):
        from dmlcloud_b200.checkpoint import CheckpointDir

        d = CheckpointDir(tmp_path / 'run')
        d.create()
        return d

    def test_state_tree_round_trips_under_every_tag(self, tmp_path):
        from dmlcloud_b200.checkpoint import AsyncSnapshot

        d = self._dir(tmp_path)
        snap = AsyncSnapshot(d, 'cpu')
        state = {'models': {'m': {'w': torch.arange(6.0).reshape(2, 3)}}, 'optimizers': {'o': {'state': [{'step': torch.tensor(7)}],
                 'groups': ({'lr': 0.1},)}}, 'stage_epoch': 4, 'tracker': {'epoch': 4, 'histories': {'x': [1.0, None]}}}
        snap.save(state, ['latest', 'epoch_3', 'best_m'])
        snap.wait()
        assert snap.written == 1
        for tag in ('latest', 'epoch_3', 'best_m'):
            assert d.has_state(tag)
            got = d.load_state(tag)
            assert torch.equal(got['models']['m']['w'], state['models']['m']['w'])
            assert int(got['optimizers']['o']['state'][0]['step']) == 7 and got['optimizers']['o']['groups'] == ({'lr': 0.1},)
            assert got['stage_epoch'] == 4 and got['tracker'] == state['tracker']
        assert not list((d.path / 'state').glob('*.tmp'))  # written next to the target, then renamed over it

    def test_a_new_save_waits_for_the_previous_write_and_overwrites_latest(self, tmp_path):
        from dmlcloud_b200.checkpoint import AsyncSnapshot

        d = self._dir(tmp_path)
        snap = AsyncSnapshot(d, 'cpu')
        gate, order = threading.Event(), []
        plain_save = d.save_state

        def slow_save(state, tag):
            if state['n'] == 1:
                gate.wait(10)
            order.append(state['n'])
            plain_save(state, tag)

        d.save_state = slow_save
        snap.save({'n': 1}, ['latest'])
        threading.Timer(0.2, gate.set).start()
        snap.save({'n': 2}, ['latest'])  # joins the first writer before it reuses the staging
        snap.wait()
        assert order == [1, 2] and d.load_state('latest')['n'] == 2 and snap.written == 2

    def test_a_failed_write_is_raised_by_the_next_wait_once(self, tmp_path):
        from dmlcloud_b200.checkpoint import AsyncSnapshot

        d = self._dir(tmp_path)
        snap = AsyncSnapshot(d, 'cpu')

        def broken(state, tag):
            raise OSError('disk full')

        d.save_state = broken
        snap.save({'n': 1}, ['latest'])
        with pytest.raises(OSError, match='disk full'):
            snap.wait()
        snap.wait()  # reported once
        assert snap.written == 0

    def test_state_files_need_a_valid_run_directory(self, tmp_path):
        from dmlcloud_b200.checkpoint import CheckpointDir

        d = CheckpointDir(tmp_path / 'missing')
        with pytest.raises(ValueError, match='not valid'):
            d.save_state({'n': 1})
        assert not d.has_state('latest')


class TestSyncBnConversionHostSide:
    """syncbn.convert (the counterpart of torch's convert_sync_batchnorm at reference pipeline.py:71) is module surgery and
    needs no GPU; neither do the cases in which the layer does not exchange anything (evaluation mode, world size 1)."""

    class _Comm:
        def __init__(self, world):
            self.world, self.rank = world, 0

    def _net(self):
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4, eps=1e-3, momentum=0.2),
                                   torch.nn.Sequential(torch.nn.ReLU(), torch.nn.BatchNorm2d(4, affine=False)),
                                   torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 5), torch.nn.BatchNorm1d(5, track_running_stats=False))

    def test_convert_keeps_parameters_buffers_keys_and_flags(self):
        from dmlcloud_b200.syncbn import PeerSyncBatchNorm, convert

        net = self._net()
        net[1].running_mean.uniform_(-1, 1)
        net.eval()
        before = dict(net.state_dict())
        comm = self._Comm(2)
        out = convert(net, comm)
        layers = [m for m in out.modules() if isinstance(m, PeerSyncBatchNorm)]
        assert len(layers) == 3 and all(m.comm is comm for m in layers)
        assert not any(type(m) in (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d) for m in out.modules())
        assert list(out.state_dict()) == list(before)
        for k, v in out.state_dict().items():
            assert v.data_ptr() == before[k].data_ptr()  # the very same storage: optimizers built before keep working
        assert (layers[0].eps, layers[0].momentum, layers[0].affine) == (1e-3, 0.2, True)
        assert layers[1].affine is False and layers[1].weight is None
        assert layers[2].track_running_stats is False and layers[2].running_mean is None
        assert not any(m.training for m in layers)
        assert convert(out, comm) is out and sum(isinstance(m, PeerSyncBatchNorm) for m in out.modules()) == 3  # idempotent

    @pytest.mark.parametrize('world,training', [(1, True), (2, False)])
    def test_without_an_exchange_the_layer_is_plain_batchnorm(self, world, training):
        import copy

        from dmlcloud_b200.syncbn import convert

        ref = self._net()[:3]
        mine = convert(copy.deepcopy(ref), self._Comm(world))
        ref.train(training)
        mine.train(training)
        x = torch.randn(8, 3, 8, 8)
        for _ in range(2):
            assert torch.equal(mine(x), ref(x))
        for (k, a), (_, b) in zip(mine.state_dict().items(), ref.state_dict().items()):
            assert torch.equal(a, b), k  # running statistics and num_batches_tracked advance alike

    def test_training_with_peers_needs_cuda_inputs(self):
        from dmlcloud_b200.syncbn import convert

        net = convert(torch.nn.BatchNorm2d(3), self._Comm(2))
        with pytest.raises(ValueError, match='CUDA'):
            net(torch.randn(2, 3, 4, 4))
        with pytest.raises(ValueError, match='at least 2D'):
            net(torch.randn(3))
