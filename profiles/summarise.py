"""Turns ncu outputs (gpurun_out/*.csv / *.ncu-rep read with `ncu -i`) into the small text summaries committed
under profiles/.  Usage: python profiles/summarise.py launches <launches.csv> | raw <raw.csv> | source <src.csv>"""
import collections
import csv
import re
import sys

RAW_KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    H, body = rows[hdr], rows[hdr + 1:]
    ki, vi = H.index('Kernel Name'), H.index('Metric Value')
    agg = collections.OrderedDict()
    for r in body:
        n = re.sub(r'\(.*', '', r[ki])[:64]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    print("%-66s %6s %12s %7s %10s" % ("kernel", "n", "total_us", "share", "avg_us"))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-66s %6d %12.1f %7.3f %10.1f" % (n, a[0], a[1] / 1e3, a[1] / tot, a[1] / a[0] / 1e3))


def raw(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        print('---', r[H.index('Kernel Name')][:90])
        for k in RAW_KEYS:
            if k in H:
                print("%-72s %20s %s" % (k, r[H.index(k)], U[H.index(k)]))
        for i, h in enumerate(H):
            if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct') and float(r[i] or 0) >= 2.0:
                print("%-72s %20s %s" % (h, r[i], U[i]))


def source(path):
    rows = list(csv.reader(open(path)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    for si in range(len(secs) - 1):
        start, end = secs[si], secs[si + 1]
        H, body = rows[start + 1], rows[start + 2:end]
        isrc, ismp, iex = H.index('Source'), H.index('# Samples'), H.index('Instructions Executed')
        tot_ex = sum(int(r[iex]) for r in body if r[iex].isdigit())
        tot_s = sum(int(r[ismp]) for r in body if r[ismp].isdigit())
        print('---', rows[start][1][:90])
        print("SASS instructions %d, executed warp-instructions %d, samples %d" % (len(body), tot_ex, tot_s))
        ops, smp = collections.Counter(), collections.Counter()
        for r in body:
            if not r[iex].isdigit():
                continue
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
            op = m.group(2).split('.')[0] if m else '?'
            ops[op] += int(r[iex])
            smp[op] += int(r[ismp])
        for op, c in ops.most_common(16):
            print("  %-10s executed %12d (%5.3f)  stall samples %7d (%5.3f)" % (op, c, c / tot_ex, smp[op], smp[op] / max(tot_s, 1)))
        cols = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
        tot = collections.Counter()
        for r in body:
            for i in cols:
                try:
                    tot[H[i]] += int(r[i])
                except ValueError:
                    pass
        print("  stall samples:", ", ".join("%s %d" % kv for kv in tot.most_common(8)))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "source": source}[sys.argv[1]](sys.argv[2])
