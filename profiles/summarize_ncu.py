"""Turns an .ncu-rep (brought back in gpurun_out/) into the small CSV summary that is committed under profiles/.
    python profiles/summarize_ncu.py gpurun_out/prof_bucket_r1.ncu-rep profiles/r1_bucket_kernels_ncu_full.csv
"""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__cycles_active.avg']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols] + ['dram_bytes_total', 'GB/s (dram bytes / duration)'])
        w.writerow([units[c] for c in cols] + ['byte', 'GB/s'])
        for r in rows[2:]:
            def val(name):
                i = hdr.index(name)
                x = float(r[i].replace(',', ''))
                u = units[i].lower()
                scale = {'gbyte': 1e9, 'mbyte': 1e6, 'kbyte': 1e3, 'byte': 1, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9,
                         'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1}.get(u, 1)
                return x * scale
            total = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
            w.writerow([r[c] for c in cols] + [f'{total:.0f}', f'{total / val("gpu__time_duration.sum") / 1e9:.1f}'])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
