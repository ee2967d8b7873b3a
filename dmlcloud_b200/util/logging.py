"""Run logging: stdout/stderr tee into the checkpoint dir, rank-aware log levels, the start-of-run banner.

Counterpart of the reference's util/logging.py (IORedirector [18-81], DevNullIO [84-90], add_log_handlers [93-108],
flush_log_handlers [111-116], experiment_header [119-128], general_diagnostics [131-173]) — host text only, nothing
here touches the device.
"""
import io
import logging
import os
import subprocess
import sys
from pathlib import Path

import torch
import torch.distributed as dist

from . import host


class _Branch:
    """A writable that forwards to the log file and to one of the original process streams."""

    def __init__(self, tee, stream_name):
        self._tee, self._name = tee, stream_name

    def write(self, text):
        if self._tee.file is not None:  # handlers created while the tee was active may outlive it
            self._tee.file.write(text)
        return getattr(self._tee, self._name).write(text)

    def flush(self):
        if self._tee.file is not None:
            self._tee.file.flush()
        getattr(self._tee, self._name).flush()

    def isatty(self):
        return False


class IORedirector:
    """install(): everything printed to stdout / stderr is also appended to `log_file`; uninstall() undoes it.
    Usable as a context manager."""

    def __init__(self, log_file: Path):
        self.path = Path(log_file)
        self.file = self.stdout = self.stderr = None

    def install(self):
        if self.file is not None:
            return
        self.file = self.path.open('a')
        self.stdout, self.stderr = sys.stdout, sys.stderr
        for s in (self.stdout, self.stderr):
            s.flush()
        sys.stdout, sys.stderr = _Branch(self, 'stdout'), _Branch(self, 'stderr')

    def uninstall(self):
        if self.file is None:
            return
        for s in (sys.stdout, sys.stderr):
            s.flush()
        sys.stdout, sys.stderr = self.stdout, self.stderr
        self.file.close()
        self.file = None

    __enter__ = lambda self: (self.install(), self)[1]  # noqa: E731

    def __exit__(self, *exc):
        self.uninstall()


class DevNullIO(io.TextIOBase):
    """Text sink (ranks other than 0 print their progress table here)."""

    def write(self, msg):
        return 0


def add_log_handlers(logger: logging.Logger):
    """INFO and below -> stdout, WARNING and above -> stderr; only rank 0 logs INFO."""
    if logger.hasHandlers():
        return
    logger.setLevel(logging.INFO if dist.get_rank() == 0 else logging.WARNING)
    plain = logging.Formatter()
    quiet = logging.StreamHandler(sys.stdout)
    quiet.setLevel(logging.DEBUG)
    quiet.addFilter(lambda rec: rec.levelno < logging.WARNING)
    loud = logging.StreamHandler()
    loud.setLevel(logging.WARNING)
    for h in (quiet, loud):
        h.setFormatter(plain)
        logger.addHandler(h)


def flush_log_handlers(logger: logging.Logger):
    for h in logger.handlers:
        h.flush()


def experiment_header(name, checkpoint_dir, date) -> str:
    rows = [f'...............  Experiment: {name or "N/A"}  ...............',
            f'- Date: {date}',
            f'- Checkpoint Dir: {checkpoint_dir or "N/A"}',
            f'- Training on {dist.get_world_size()} GPUs']
    return '\n'.join(rows) + '\n'


def _section(title, items):
    return [f'* {title}:'] + [f'    - {k}: {v}' if k is not None else f'    - {v}' for k, v in items]


def general_diagnostics() -> str:
    import dmlcloud_b200

    env = os.environ
    out = _section('GENERAL', [
        ('argv', sys.argv), ('cwd', Path.cwd()), ('host (root)', env.get('HOSTNAME')), ('user', env.get('USER')),
        ('git-hash', host.git_revision()), ('conda-env', env.get('CONDA_DEFAULT_ENV', 'N/A')),
        ('sys-prefix', sys.prefix), ('backend', dist.get_backend()), ('cuda', torch.cuda.is_available()),
    ])
    if torch.cuda.is_available():
        try:
            listing = subprocess.run(['nvidia-smi', '-L'], capture_output=True, text=True).stdout.splitlines()
        except OSError:
            listing = []
        out += _section('GPUs (root)', [(None, line) for line in listing])
        try:
            lib = dmlcloud_b200._native.device_info(torch.cuda.current_device())
            out += _section('libdmlb', [('SMs', lib['sm_count']), ('L2 bytes', lib['l2_bytes']), ('cc', lib['cc'])])
        except Exception as exc:  # noqa: BLE001 - diagnostics must never stop a run
            out += _section('libdmlb', [('unavailable', exc)])
    versions = [('python', sys.version), ('dmlcloud_b200', dmlcloud_b200.__version__), ('cuda', torch.version.cuda)]
    try:
        versions.append(('driver', Path('/proc/driver/nvidia/version').read_text().splitlines()[0]))
    except (OSError, IndexError):
        pass
    out += _section('VERSIONS', versions + list(host.package_versions().items()))
    slurm = host.slurm_environment()
    if slurm:
        out += _section('SLURM', list(slurm.items()))
    return '\n'.join(out) + '\n'


def run_banner(pipeline) -> str:
    """Everything `TrainingPipeline._pre_run` logs once at start: diagnostics, per-rank devices, the config."""
    from .config import Conf
    from .distributed import all_gather_object

    text = general_diagnostics()
    devices = all_gather_object(str(pipeline.device))
    text += '\n'.join(['* DEVICES:'] + [f'    - [Rank {i}] {d}' for i, d in enumerate(devices)]) + '\n'
    text += '\n'.join(['* CONFIG:'] + [f'    {line}' for line in Conf.to_yaml(pipeline.config, resolve=True).splitlines()])
    return text
