"""Per-step kernel breakdown from an `ncu --metrics gpu__time_duration.sum` launch list of bench.py.
    python profiles/summarize_launches.py gpurun_out/launches.csv profiles/rN_launches.md [steps_to_average] [delimiter]
Steps are delimited by the fused step exchange (dmlb::allreduce_oneshot_kernel / allreduce_twoshot_kernel: exactly one per
captured step since round 2; round 1's lists were delimited by dmlb::metric_reduce_kernel).  ncu times are cold-cache and
serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import sys


def main(src, dst, last=15, delimiter='allreduce_'):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    names = [r[ki] for r in data]
    vals = []
    for r in data:
        v = float(r[vi].replace(',', ''))
        vals.append(v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v))
    marks = [i for i, n in enumerate(names) if delimiter in n]
    segs = [(marks[i - 1] + 1, marks[i] + 1) for i in range(max(1, len(marks) - last), len(marks))]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for a, b in segs:
        for j in range(a, b):
            agg[names[j]][0] += 1
            agg[names[j]][1] += vals[j]
    ns = len(segs)
    tot = sum(v[1] for v in agg.values()) / ns
    ours = sum(v[1] for k, v in agg.items() if 'dmlb::' in k) / ns
    with open(dst, 'w') as f:
        f.write(f'# Launch list summary: {src}\n\n')
        f.write(f'{len(data)} launches captured; averaged over the last {ns} steps: '
                f'**{sum(v[0] for v in agg.values()) / ns:.1f} launches/step, {tot:.1f} us GPU time/step** '
                f'(ncu: cold-cache, serialised).\n\n')
        f.write(f'libdmlb kernels: {sum(v[0] for k, v in agg.items() if "dmlb::" in k) / ns:.1f} launches/step, '
                f'{ours:.1f} us/step = **{100 * ours / tot:.1f} % of the step\'s GPU time**; the rest is the user model '
                f'(cuDNN / ATen kernels of the model under bf16 autocast).\n\n')
        f.write('| launches/step | us each | us/step | share | kernel |\n|---:|---:|---:|---:|---|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| {v[0] / ns:.1f} | {v[1] / v[0]:.2f} | {v[1] / ns:.1f} | {100 * v[1] / ns / tot:.1f}% | `{k[:110]}` |\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 15,
         sys.argv[4] if len(sys.argv) > 4 else 'allreduce_')
