"""Turns gpurun_out/ ncu captures into the tracked summaries under profiles/ (run here, no GPU needed).
usage: summarise_profiles.py launches <csv> <out.txt> <title> | full <ncu-rep> <out.csv>"""
import collections, csv, subprocess, sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']


def launches(path, out, title):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == 'ID':
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'us': 1e3, 'ms': 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki].split('(')[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["# " + title, "# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
             "%-48s %6s %14s %7s %12s" % ("kernel", "count", "total_ns", "share", "avg_ns")]
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("%-48s %6d %14.0f %6.1f%% %12.0f" % (n[:48], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
    open(out, 'w').write("\n".join(lines) + "\n")
    print("\n".join(lines))


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(open(out, 'w'))
    w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(rows) - 2)])
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([hdr[i], units[i]] + [r[i][:60] for r in rows[2:]])
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
