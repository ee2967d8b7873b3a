"""Host logic of PeerComm's NVSwitch-multicast arena setup (dmlcloud_b200/gradsync.py `_setup_multicast`), on CPU with gloo.

The CUDA driver calls (cuMemCreate / export / import / cuMulticast*) are behind libdmlb's `dmlb_vmm_*` / `dmlb_mc_*` entry
points; here a fake library stands in for them — arenas are memfd files, "importing" a descriptor reads the owner tag
written into it — so that what is tested is what Python does around them: the votes (all ranks succeed or all raise), the
descriptor exchange over unix sockets (every rank ends up with every peer's arena, rank 0's multicast object reaches all),
and that no socket path or descriptor is left behind, also when one rank fails half-way.  The real thing runs in
tests/test_gpu_gradsync.py::test_nvls_allreduce_two_gpus on a box with NVSwitch.
"""
import ctypes
import glob
import json
import os
import tempfile
from pathlib import Path

import pytest
import torch

from helpers import init_gloo, spawn

OK = 0


class FakeVmmLib:
    """dmlb_vmm_* / dmlb_mc_* of include/dmlb.h on memfd files.  Pointers and handles are small integers that encode the
    owner: ptr = 0x1000 * (1 + owner rank), multicast handle = 0x77."""

    def __init__(self, rank, fail_at=None):
        self.rank, self.fail_at, self.calls, self.open_fds = rank, fail_at, [], set()

    def _fail(self, name):
        self.calls.append(name)
        return -1 if self.fail_at == name else OK

    def _memfd(self, tag):
        fd = os.memfd_create(tag)
        os.write(fd, tag.encode())
        return fd

    def dmlb_vmm_granularity(self, dev, world):
        return 2 << 20

    def dmlb_vmm_alloc(self, dev, size, ptr, fd, handle):
        if self._fail('vmm_alloc'):
            return -1
        assert size % (2 << 20) == 0
        ptr._obj.value, fd._obj.value, handle._obj.value = 0x1000 * (1 + self.rank), self._memfd(f'arena{self.rank}'), 100 + self.rank
        return OK

    def dmlb_mc_create(self, world, size, fd, handle):
        if self._fail('mc_create'):
            return -1
        fd._obj.value, handle._obj.value = self._memfd('mc'), 0x77
        return OK

    def dmlb_vmm_import(self, dev, fd, size, ptr, handle):
        if self._fail('vmm_import'):
            return -1
        tag = os.pread(fd, 16, 0).decode()
        assert tag.startswith('arena'), tag
        owner = int(tag[5:])
        ptr._obj.value, handle._obj.value = 0x1000 * (1 + owner), 100 + owner
        return OK

    def dmlb_mc_import(self, fd, handle):
        if self._fail('mc_import'):
            return -1
        assert os.pread(fd, 16, 0).decode() == 'mc'
        handle._obj.value = 0x77
        return OK

    def dmlb_mc_add_device(self, handle, dev):
        assert handle == 0x77
        return self._fail('mc_add_device')

    def dmlb_mc_bind(self, handle, dev, own_handle, size, mc_ptr):
        if self._fail('mc_bind'):
            return -1
        assert handle == 0x77 and own_handle == 100 + self.rank
        mc_ptr._obj.value = 0xABC000
        return OK


def _open_fds():
    return set(os.listdir('/proc/self/fd'))


def _worker(rank, world, initfile, outdir, failing_rank, fail_at):
    init_gloo(rank, world, initfile)
    import torch.distributed as dist

    from dmlcloud_b200 import gradsync
    from dmlcloud_b200.gradsync import PeerComm

    gradsync._device_uuid = lambda index: f'GPU-{index}'  # one (pretend) GPU per rank
    comm = object.__new__(PeerComm)
    comm.device, comm.group, comm.world, comm.rank = torch.device('cuda', rank), None, world, rank
    comm.arena_bytes, comm._vmm, comm._own = 9 << 20, None, None
    lib = FakeVmmLib(rank, fail_at if rank == failing_rank else None)
    before = _open_fds()
    result = {'raised': None}
    try:
        arenas = comm._setup_multicast(lib)
        result['arenas'] = [int(a) for a in arenas]
        result['bytes'] = comm._vmm['bytes']
        result['mc_handle'], result['mc_ptr'] = comm._vmm['mc_handle'], comm._vmm['mc_ptr'].value
        result['peers'] = sorted(int(p.value) for p, _ in comm._vmm['peers'])
    except RuntimeError as exc:
        result['raised'] = str(exc)
    result['leaked_fds'] = len(_open_fds() - before)
    result['calls'] = lib.calls
    dist.barrier()
    result['socket_paths_left'] = len(glob.glob(os.path.join(tempfile.gettempdir(), 'dmlb_*_%d.sock' % rank)))
    Path(outdir, f'r{rank}.json').write_text(json.dumps(result))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_every_rank_maps_every_arena_and_the_multicast_object(world):
    out = spawn(_worker, world, -1, None, timeout=120)
    for r in range(world):
        res = json.loads((out / f'r{r}.json').read_text())
        assert res['raised'] is None, res
        assert res['arenas'] == [0x1000 * (1 + q) for q in range(world)]  # arena of rank q at index q, own included
        assert res['peers'] == [0x1000 * (1 + q) for q in range(world) if q != r]
        assert res['bytes'] == 10 << 20  # 9 MiB rounded up to the 2 MiB granularity
        assert res['mc_handle'] == 0x77 and res['mc_ptr'] == 0xABC000
        assert res['leaked_fds'] == 0 and res['socket_paths_left'] == 0, res
        assert res['calls'].count('vmm_import') == world - 1 and res['calls'].count('mc_bind') == 1
        assert ('mc_create' in res['calls']) == (r == 0) and ('mc_import' in res['calls']) == (r != 0)


@pytest.mark.parametrize('fail_at', ['vmm_alloc', 'vmm_import', 'mc_add_device', 'mc_bind'])
def test_one_failing_rank_makes_every_rank_raise_together(fail_at):
    """All-or-none: a local failure is reported through the next vote, after the rank has still taken part in every
    collective and socket exchange before it — no rank is left waiting, none ends up with a communicator."""
    out = spawn(_worker, 2, 1, fail_at, timeout=120)
    for r in range(2):
        res = json.loads((out / f'r{r}.json').read_text())
        assert res['raised'] is not None and 'failed on ranks [1]' in res['raised'], res
        assert res['leaked_fds'] == 0 and res['socket_paths_left'] == 0, res
