"""`import dmlcloud` for code written against the reference package.

    import dmlcloud_b200.compat            # registers the alias (idempotent), then:
    from dmlcloud.pipeline import TrainingPipeline
    from dmlcloud.metrics import MetricTracker, Reduction

Every reference module name on the drop-in boundary resolves to its dmlcloud_b200 counterpart:
dmlcloud, dmlcloud.stage, dmlcloud.pipeline, dmlcloud.metrics, dmlcloud.checkpoint, dmlcloud.util,
dmlcloud.util.distributed, dmlcloud.util.data, dmlcloud.util.logging.  The alias is refused when a real `dmlcloud`
package is already imported (the installed reference under oracle/_ref, for instance) — the two must never mix.

    python -m dmlcloud_b200.compat -m pytest path/to/reference/test/test_metrics.py     # run a module under the alias
"""
import importlib
import runpy
import sys

_MODULES = ('stage', 'pipeline', 'metrics', 'checkpoint', 'util', 'util.distributed', 'util.data', 'util.logging')


def install():
    import dmlcloud_b200

    existing = sys.modules.get('dmlcloud')
    if existing is not None and existing is not dmlcloud_b200:
        raise ImportError('a different `dmlcloud` package is already imported; refusing to alias over it')
    sys.modules['dmlcloud'] = dmlcloud_b200
    for name in _MODULES:
        sys.modules[f'dmlcloud.{name}'] = importlib.import_module(f'dmlcloud_b200.{name}')
    return dmlcloud_b200


install()

if __name__ == '__main__':
    if len(sys.argv) >= 3 and sys.argv[1] == '-m':
        module, sys.argv = sys.argv[2], sys.argv[2:]
        runpy.run_module(module, run_name='__main__', alter_sys=True)
    else:
        raise SystemExit('usage: python -m dmlcloud_b200.compat -m <module> [args...]')
