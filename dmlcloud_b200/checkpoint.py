"""Checkpoint directory: `<root>/<name>-<Y.m.d-H.M>-<id>/` holding `.dmlcloud` (marker), `log.txt`, `config.yaml`,
`.slurm-jobid` and — new — `state/*.pt`.

Public names and on-disk layout follow the reference's dmlcloud/checkpoint.py (sanitize_filename [12], generate_id [16],
generate_checkpoint_path [21-34], find_slurm_checkpoint [37-48], CheckpointDir [51-123]) so existing run directories
resume.  The reference never stores training state (register_model's save_* arguments are ignored, pipeline.py:61-64);
`save_state` / `load_state` below do: model / optimizer / scheduler / MetricTracker state, written
atomically, so a resumed run continues with bit-identical `tracker.epoch` and `stage.current_epoch`.
"""
import logging
import secrets
from datetime import datetime
from pathlib import Path
from typing import Optional

from .util.config import Conf
from .util.host import slurm_job_id

MARKER, LOG, CONFIG, JOBID, STATE = '.dmlcloud', 'log.txt', 'config.yaml', '.slurm-jobid', 'state'


def sanitize_filename(filename: str) -> str:
    return filename.replace('/', '_')


def generate_id() -> str:
    token = secrets.token_urlsafe(5)
    return token.translate(str.maketrans('-_', 'ab'))


def generate_checkpoint_path(root, name: Optional[str] = None, creation_time: Optional[datetime] = None) -> Path:
    # The stamp is "now" (the reference ignores `creation_time` too, checkpoint.py:29-32); kept for signature parity.
    label = sanitize_filename(name if name is not None else 'run')
    return Path(root) / '-'.join((label, datetime.now().strftime('%Y.%m.%d-%H.%M'), generate_id()))


def find_slurm_checkpoint(root) -> Optional[Path]:
    """The run directory under `root` that was created by the current SLURM job (requeue -> resume), if any."""
    job = slurm_job_id()
    if job is None:
        return None
    hits = (p for p in Path(root).iterdir() if CheckpointDir(p).is_valid and CheckpointDir(p).slurm_job_id == job)
    return next(hits, None)


class _Member:
    """Path of a fixed-name file inside the directory, exposed as a read-only attribute."""

    def __init__(self, filename):
        self.filename = filename

    def __get__(self, obj, owner=None):
        return self if obj is None else obj.path / self.filename


class CheckpointDir:
    config_file = _Member(CONFIG)
    indicator_file = _Member(MARKER)
    log_file = _Member(LOG)
    slurm_file = _Member(JOBID)
    state_dir = _Member(STATE)

    def __init__(self, path):
        self.path = Path(path).resolve()
        self.logger = logging.getLogger('dmlcloud')

    def __str__(self):
        return str(self.path)

    def __repr__(self):
        return f'CheckpointDir({self.path})'

    @property
    def exists(self) -> bool:
        return self.path.exists()

    @property
    def is_valid(self) -> bool:
        return self.path.is_dir() and self.indicator_file.exists()

    @property
    def slurm_job_id(self) -> Optional[str]:
        return self.slurm_file.read_text() if self.slurm_file.exists() else None

    def create(self):
        if self.exists:
            raise ValueError(f'Checkpoint directory already exists: {self.path}')
        self.path.mkdir(parents=True)
        for member in (self.indicator_file, self.log_file):
            member.touch()
        job = slurm_job_id()
        if job is not None:
            self.slurm_file.write_text(job)

    def save_config(self, config):
        if not self.exists:
            raise ValueError(f'Checkpoint directory does not exist: {self.path}')
        with self.config_file.open('w') as fh:
            Conf.save(config, fh)

    def load_config(self):
        if not self.is_valid:
            raise ValueError(f'Checkpoint directory is not valid: {self.path}')
        with self.config_file.open() as fh:
            return Conf.load(fh)

    # ---- training state (new) ----------------------------------------------------------------------------------------
    def _state_path(self, tag):
        return self.state_dir / f'{sanitize_filename(tag)}.pt'

    def save_state(self, state: dict, tag: str = 'latest'):
        import torch

        if not self.is_valid:
            raise ValueError(f'Checkpoint directory is not valid: {self.path}')
        self.state_dir.mkdir(exist_ok=True)
        final = self._state_path(tag)
        scratch = final.with_suffix('.pt.tmp')
        torch.save(state, scratch)
        scratch.replace(final)  # rename is atomic: a crash mid-write never leaves a torn snapshot

    def has_state(self, tag: str = 'latest') -> bool:
        return self._state_path(tag).exists()

    def load_state(self, tag: str = 'latest', map_location='cpu'):
        import torch

        return torch.load(self._state_path(tag), map_location=map_location, weights_only=False)


class AsyncSnapshot:
    """Asynchronous state snapshots: device -> staging arena (same stream as the training step) -> pinned
    host memory (side stream) -> file (writer thread).  One snapshot is in flight at a time; the staging buffers are
    reused, so a new `save` first waits for the previous write (normally long finished)."""

    def __init__(self, checkpoint_dir, device):
        import threading

        import torch

        self.dir = checkpoint_dir
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == 'cuda' else None
        self._staging = {}  # path in the state tree -> (device copy, pinned host copy)
        self._thread = None
        self._error = None
        self._lock = threading.Lock()
        self.written = 0

    def _stage(self, obj, path, copies):
        import torch

        if isinstance(obj, torch.Tensor):
            if not obj.is_cuda:
                return obj
            slot = self._staging.get(path)
            if slot is None or slot[0].shape != obj.shape or slot[0].dtype != obj.dtype:
                slot = (torch.empty_like(obj, memory_format=torch.contiguous_format),
                        torch.empty(obj.shape, dtype=obj.dtype, pin_memory=True))
                self._staging[path] = slot
            slot[0].copy_(obj, non_blocking=True)  # D2D on the caller's stream: the original may change right after
            copies.append(slot)
            return slot[1]
        if isinstance(obj, dict):
            return {k: self._stage(v, f'{path}/{k}', copies) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            out = [self._stage(v, f'{path}/{i}', copies) for i, v in enumerate(obj)]
            return out if isinstance(obj, list) else tuple(out)
        return obj

    def save(self, state, tags):
        import threading

        import torch

        self.wait()
        copies = []
        host_state = self._stage(state, '', copies)
        event = None
        if self.stream is not None and copies:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for dev_copy, host_copy in copies:
                    host_copy.copy_(dev_copy, non_blocking=True)
                event = torch.cuda.Event()
                event.record()

        def write():
            try:
                if event is not None:
                    event.synchronize()
                for tag in tags:
                    self.dir.save_state(host_state, tag)
                self.written += 1
            except Exception as exc:  # noqa: BLE001 - surfaced by the next wait()
                self._error = exc

        self._thread = threading.Thread(target=write, name='dmlcloud-snapshot', daemon=True)
        self._thread.start()

    def wait(self):
        t, self._thread = self._thread, None
        if t is not None:
            t.join()
        if self._error is not None:
            err, self._error = self._error, None
            raise err
