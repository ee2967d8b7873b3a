"""Process-group bootstrap + small object collectives (host control plane, no device arithmetic).

Public names follow the reference's dmlcloud/util/distributed.py so user scripts and test fixtures keep working:
init_process_group_{dummy,slurm,MPI,auto}, deinitialize_torch_distributed, is_root, root_only, root_first,
rank / world_size / local_rank / local_world_size / local_node, print_worker, print_root,
all_gather_object, gather_object, broadcast_object  (reference util/distributed.py:39-259).

Built differently: one `Placement` record describes where this process sits (filled by whichever launcher started it),
and every init function is `_join(placement, how-to-rendezvous)`.  Deliberate fix: under torchrun the
reference leaves the placement empty (its env:// branch, [237-238]), so all ranks share cuda:0; here torchrun's
RANK / WORLD_SIZE / LOCAL_RANK / LOCAL_WORLD_SIZE / GROUP_RANK are read — one process per GPU on the 8xB200 box.
"""
import functools
import os
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from .host import find_free_port, local_ips

DEFAULT_PORT = os.environ.get('DMLCLOUD_PORT', 41312)  # "dml" on a phone keypad


@dataclass
class Placement:
    how: Optional[str] = None
    rank: Optional[int] = None
    world: Optional[int] = None
    local_rank: Optional[int] = None
    local_world: Optional[int] = None
    node: Optional[int] = None


_here = Placement()


def _placement_getter(field):
    def getter():
        return getattr(_here, field)

    getter.__name__ = field
    return getter


rank = _placement_getter('rank')
world_size = _placement_getter('world')
local_rank = _placement_getter('local_rank')
local_world_size = _placement_getter('local_world')
local_node = _placement_getter('node')


def has_slurm():
    return 'SLURM_PROCID' in os.environ


def has_environment():
    return 'MASTER_PORT' in os.environ


def has_mpi():
    from .host import try_import

    return try_import('mpi4py') is not None


def is_root():
    return dist.get_rank() == 0


def root_only(fn):
    """Decorator: the wrapped callable runs on rank 0 and is a no-op (returning None) elsewhere."""

    @functools.wraps(fn)
    def on_root(*args, **kwargs):
        return fn(*args, **kwargs) if is_root() else None

    return on_root


@contextmanager
def root_first():
    """`with root_first():` — rank 0 executes the body, then everybody else does (dataset downloads etc.)."""
    if not is_root():
        dist.barrier()
    try:
        yield
    finally:
        if is_root():
            dist.barrier()


def print_worker(msg, barrier=True, flush=True):
    tag = f'Worker {rank()}' + (f'({local_node()}.{local_rank()})' if local_node() is not None else '')
    if barrier:
        dist.barrier()
    print(f'{tag}:{msg}', flush=flush)
    if barrier:
        dist.barrier()


def print_root(msg, flush=True):
    if is_root():
        print(msg, flush=flush)


# ---- object collectives (pickle based, setup-time only) --------------------------------------------------------------
def all_gather_object(obj, group=None):
    slots = [None] * dist.get_world_size(group)
    dist.all_gather_object(slots, obj, group=group)
    return slots


def gather_object(obj, dst=0, group=None):
    slots = [None] * dist.get_world_size(group) if dist.get_rank() == dst else None
    dist.gather_object(obj, slots, dst=dst, group=group)
    return slots


def broadcast_object(obj, src=0, group=None, device=None):
    # `device` is accepted for signature compatibility; pickled objects always travel through the CPU path
    cell = [obj]
    dist.broadcast_object_list(cell, src=src, group=group)
    return cell[0]


# ---- joining a process group -----------------------------------------------------------------------------------------
def _mixed_backend():
    """gloo for CPU tensors (barriers, pickles), NCCL for CUDA tensors (gradient / metric exchange over NVLink)."""
    if dist.is_nccl_available() and torch.cuda.is_available():
        return 'cpu:gloo,cuda:nccl'
    return 'gloo'


def _join(placement, **init_kwargs):
    global _here
    _here = placement
    if torch.cuda.is_available() and placement.local_rank is not None:
        torch.cuda.set_device(placement.local_rank % torch.cuda.device_count())
    dist.init_process_group(**init_kwargs)


def init_process_group_dummy(**kwargs):
    """World of one over an in-memory HashStore: single-GPU runs and unit tests."""
    backend = kwargs.pop('backend', None) or _mixed_backend()
    _join(Placement('dummy', 0, 1, 0, 1, 0), store=dist.HashStore(), rank=0, world_size=1, backend=backend, **kwargs)


def init_process_group_slurm(port=DEFAULT_PORT, **kwargs):
    e = os.environ
    here = Placement('slurm', int(e['SLURM_PROCID']), int(e['SLURM_NTASKS']), int(e['SLURM_LOCALID']),
                     int(e['SLURM_STEP_TASKS_PER_NODE']), int(e['SLURM_NODEID']))
    _join(here, init_method=f'tcp://{e["SLURM_SRUN_COMM_HOST"]}:{port}', world_size=here.world, rank=here.rank,
          **kwargs)


def init_process_group_MPI(ip_idx=0, port=DEFAULT_PORT, **kwargs):
    """Rendezvous address travels over mpi4py (works even when torch was built without MPI): rank 0 publishes
    `ip:port`, everyone connects over TCP.  port=None picks a free one; ip_idx selects among rank 0's addresses."""
    from mpi4py import MPI

    everyone = MPI.COMM_WORLD
    same_host = everyone.Split_type(MPI.COMM_TYPE_SHARED, 0, MPI.INFO_NULL)
    here = Placement('mpi', everyone.Get_rank(), everyone.Get_size(), same_host.Get_rank(), same_host.Get_size())
    address = None
    if here.rank == 0:
        address = (local_ips()[ip_idx], find_free_port() if port is None else port)
    ip, port = everyone.bcast(address, root=0)
    everyone.Barrier()
    _join(here, init_method=f'tcp://{ip}:{port}', world_size=here.world, rank=here.rank, **kwargs)


def init_process_group_env(**kwargs):
    """torchrun (`env://`): one process per GPU, placement taken from the launcher's environment."""
    e = os.environ
    r, w = int(e.get('RANK', 0)), int(e.get('WORLD_SIZE', 1))
    here = Placement('env', r, w, int(e.get('LOCAL_RANK', r)), int(e.get('LOCAL_WORLD_SIZE', w)),
                     int(e.get('GROUP_RANK', 0)))
    kwargs.setdefault('backend', _mixed_backend())
    _join(here, init_method='env://', **kwargs)


def init_process_group_auto(verbose=True, **kwargs):
    """First match wins: MASTER_PORT set -> env://; srun -> SLURM; mpi4py importable -> MPI; else a world of one."""
    for applies, init in ((has_environment, init_process_group_env), (has_slurm, init_process_group_slurm),
                          (has_mpi, init_process_group_MPI)):
        if applies():
            return init(**kwargs)
    return init_process_group_dummy()


def mpi_local_comm():
    from .host import try_import

    mpi4py = try_import('mpi4py.MPI')
    return None if mpi4py is None else mpi4py.COMM_WORLD.Split_type(mpi4py.COMM_TYPE_SHARED, 0, mpi4py.INFO_NULL)


def deinitialize_torch_distributed():
    global _here
    _here = Placement()
    dist.destroy_process_group()
