"""K5 (dmlb_adam_step_f32) roofline microbench: four 256 MiB fp32 arrays (1 GiB working set, cold for the 126 MB L2),
CUDA-event timed per launch on the launching stream; 28 algorithmic bytes per element (16 read + 12 written).
Also torch's fused Adam on the same sizes, as context (library kernel, same traffic).

    python profiles/run_adam_kernel.py > gpurun_out/adam_kernel.json
"""
import json
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from dmlcloud_b200 import _native as N  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    lib, st = N.cuda_lib(0), N.stream_ptr()
    peak = 6575.1
    try:
        peak = json.load(open(Path(__file__).resolve().parent.parent / 'MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:  # noqa: BLE001 - the fallback is the number this pool measured
        pass
    n = 1 << 26
    p, g, m, v = (torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(4))
    p.normal_()
    g.normal_()
    state = torch.zeros(2, dtype=torch.int64, device=dev)

    def timed(fn, reps=12, warm=3):
        for _ in range(warm):
            fn()
        out = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            out.append(a.elapsed_time(b) * 1e3)
        return out

    ours = timed(lambda: N.check(lib.dmlb_adam_step_f32(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3,
                                                        0.9, 0.999, 1e-8, 0.0, 0, 0, None, 0.0, state.data_ptr(), 1, None,
                                                        0, st)))
    q = torch.nn.Parameter(torch.empty(n, dtype=torch.float32, device=dev).normal_())
    q.grad = g
    ref = torch.optim.Adam([q], lr=1e-3, fused=True)
    theirs = timed(lambda: ref.step())

    def entry(name, us):
        mean = statistics.mean(us)
        gbps = n * 28 / (mean * 1e-6) / 1e9
        return {'kernel': name, 'elements': n, 'algorithmic_bytes_per_launch': n * 28, 'mean_us': round(mean, 2),
                'best_us': round(min(us), 2), 'achieved_GBps': round(gbps, 1), 'peak_GBps': peak,
                'frac': round(gbps / peak, 4)}

    print(json.dumps({'gpu': torch.cuda.get_device_name(0), 'steps_taken': int(state[0].item()),
                      'results': [entry('dmlb_adam_step_f32 (K5)', ours),
                                  entry('torch.optim.Adam(fused=True).step() (library, context)', theirs)]}, indent=1))


if __name__ == '__main__':
    main()
