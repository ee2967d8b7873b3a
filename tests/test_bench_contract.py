"""bench.py's output contract, checked without a GPU: the reference arm (`--impl reference`: the installed, unmodified
reference from oracle/_ref on the host cores — or the restatement oracle/ref_port.py where _ref did not travel) is run for
a few steps and its single JSON line validated; the native arm must refuse to run without CUDA (no CPU
fallback); the committed native bench line of this round carries every key the contract names."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e'}


def run_bench(*args, timeout=280):
    return subprocess.run([sys.executable, str(ROOT / 'bench.py'), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=str(ROOT))


def test_reference_arm_prints_one_contract_line():
    proc = run_bench('--impl', 'reference', '--steps', '3', '--warmup', '3', '--min-seconds', '0.3')
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # ONE JSON line on stdout, everything else goes to stderr
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and BASE_KEYS <= set(d)
    assert d['n_gpus'] == 1 and d['steps'] == 3 and d['warmup'] == 3
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['data'] == 'synthetic' and 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    installed = (ROOT / 'oracle' / '_ref' / 'dmlcloud' / 'stage.py').exists()
    assert cb['kind'] == ('reference' if installed else 'port') and cb['cores'] >= 1 and cb['sample']
    assert cb['value'] == d['value']
    # the stock behaviour (metrics cross ranks once per epoch) is the headline; the per-step-reduce variant sits beside it
    assert d['per_step_reduce_value'] > 0 and cb['epoch_reduce_ms'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # value is what the line says it is: samples per second over the timed steps
    assert abs(d['value'] - 32 * 1000.0 / d['ms_per_step']) <= 0.01 * d['value']


@pytest.mark.skipif(torch.cuda.is_available(), reason='CUDA present')
def test_native_arm_refuses_without_cuda():
    proc = run_bench('--steps', '2', '--warmup', '3', '--no-micro', '--no-cpu-baseline', timeout=120)
    assert proc.returncode != 0 and 'no CPU fallback' in (proc.stderr + proc.stdout)
    assert not proc.stdout.strip()  # no bench line without a measurement


def test_committed_native_line_carries_the_contract_keys():
    d = json.loads((ROOT / 'profiles' / 'r2_bench_mnist_n1.json').read_text())
    assert BASE_KEYS | {'clocks', 'gpu_launches', 'roofline', 'cpu_baseline', 'windows_ms', 'replicas_identical'} <= set(d)
    assert d['replicas_identical'] is True and len(d['windows_ms']) >= 5 and len(d['e2e']['windows_ms']) == len(d['windows_ms'])
    # `value` is the MEDIAN window, and it is not faster than its own fastest window nor slower than its slowest
    per_step = [w / d['steps'] for w in d['windows_ms']]
    assert min(per_step) - 1e-4 <= d['ms_per_step'] <= max(per_step) + 1e-4
    assert d['e2e']['value'] <= d['value'] * 1.02  # the host-fed loop does strictly more work
    assert d['cpu_baseline']['kind'] == 'reference' and d['config']['kernels_per_step'] == 2
    assert d['gpu_launches'] > 0 and d['n_gpus'] == 1
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(d['e2e'])
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    r = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r) and r['bound'] == 'hbm'
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-3 and 0.7 <= r['frac'] <= 1.05
    assert r['traffic'] is None or 0.9 <= r['traffic'] / r['algorithmic_bytes_per_launch'] <= 1.2  # no wasted re-reads
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(d['cpu_baseline'])
