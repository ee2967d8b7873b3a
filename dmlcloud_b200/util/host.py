"""Host-side odds and ends the pipeline's banner, checkpoint lookup and MPI bootstrap need.

The reference spreads these over util/{tcp,slurm,git,project,thirdparty,seed,argparse,wandb}.py; none of them is on the
data-parallel hot path (out of scope), so only what dmlcloud_b200 itself calls lives here.
"""
import importlib
import os
import random
import socket
import subprocess
import sys
from pathlib import Path

TRACKED_PACKAGES = ('torch', 'torchvision', 'numpy', 'einops', 'pandas', 'sklearn')


# ---- scheduler environment ----------------------------------------------------------------------------------------
def slurm_job_id():
    return os.environ.get('SLURM_JOB_ID')


def slurm_environment():
    keys = ('SLURM_JOB_ID', 'SLURM_STEP_ID', 'SLURM_STEP_NODELIST', 'SLURM_TASKS_PER_NODE', 'SLURM_STEP_GPUS',
            'SLURM_GPUS_ON_NODE', 'SLURM_CPUS_PER_TASK')
    return {k: os.environ.get(k) for k in keys} if slurm_job_id() is not None else {}


# ---- network --------------------------------------------------------------------------------------------------------
def find_free_port():
    with socket.socket() as probe:
        probe.bind(('', 0))
        return probe.getsockname()[1]


def local_ips():
    try:
        out = subprocess.run(['hostname', '-I'], capture_output=True, text=True, check=True).stdout.split()
        if out:
            return out
    except (OSError, subprocess.CalledProcessError):
        pass
    return socket.gethostbyname_ex(socket.gethostname())[2]


# ---- provenance -----------------------------------------------------------------------------------------------------
def launch_dir():
    main = sys.modules.get('__main__')
    path = getattr(main, '__file__', None)
    return Path(path).resolve().parent if path else Path.cwd()


def git_revision(short=False):
    cmd = ['git', 'rev-parse'] + (['--short'] if short else []) + ['HEAD']
    try:
        proc = subprocess.run(cmd, cwd=launch_dir(), capture_output=True, text=True)
    except OSError:
        return None
    return proc.stdout.strip() if proc.returncode == 0 else None


def package_versions(names=TRACKED_PACKAGES):
    """Versions of the packages that are already imported (never triggers an import)."""
    found = {}
    for name in names:
        mod = sys.modules.get(name)
        if mod is not None:
            found[name] = str(getattr(mod, '__version__', '?'))
    return found


def try_import(name):
    try:
        return importlib.import_module(name)
    except ImportError:
        return None


# ---- determinism ----------------------------------------------------------------------------------------------------
def seed_all(seed: int):
    import numpy as np
    import torch

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
