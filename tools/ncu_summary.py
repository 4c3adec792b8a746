"""Summaries of ncu outputs (launch list CSV, raw page of a .ncu-rep) for profiles/."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:70]
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"{'kernel':72s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}"]
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{n:72s} {a[0]:5d} {a[1]:11.1f} {a[1] / a[0]:9.1f} {a[1] / tot * 100:5.1f}%")
    return "\n".join(out)


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum',
        'lts__t_bytes.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']


def raw(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ni = hdr.index('Kernel Name')
    idx = [i for i, h in enumerate(hdr) if h in WANT]
    out = []
    for r in data:
        out.append(r[ni].split('(')[0][:90])
        for i in idx:
            out.append(f"    {hdr[i]:70s} {r[i]:>16s} {units[i]}")
    return "\n".join(out)


if __name__ == "__main__":
    print(launches(sys.argv[2]) if sys.argv[1] == "launches" else raw(sys.argv[2]))
