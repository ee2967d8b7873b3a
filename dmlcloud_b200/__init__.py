"""dmlcloud_b200 — the B200-native data-parallel hot path of dmlcloud, behind dmlcloud's own API.

    from dmlcloud_b200 import Stage, TrainValStage                 (reference: dmlcloud/__init__.py)
    from dmlcloud_b200.pipeline import TrainingPipeline            (reference: dmlcloud/pipeline.py)
    from dmlcloud_b200.metrics import MetricTracker, Reduction     (reference: dmlcloud/metrics.py)
    from dmlcloud_b200.util.distributed import init_process_group_auto

Importing the package never touches CUDA and never needs a GPU: the kernels live in csrc/libdmlb.so (C ABI in
include/dmlb.h), loaded through ctypes on first use.  Using a reduced metric or a DDP model without CUDA raises —
there is no CPU implementation to fall back to.
"""
__version__ = '0.1.0'

from . import _native  # noqa: E402,F401  (ctypes signatures only; does not load the library)
from .stage import Stage, TrainValStage  # noqa: E402

__all__ = ['Stage', 'TrainValStage', '__version__']
