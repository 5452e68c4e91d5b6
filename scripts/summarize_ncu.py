"""Turn gpurun_out/*.ncu-rep and launch lists into the small text summaries kept under profiles/."""
import csv, subprocess, sys, collections, io
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__cycles_active.avg', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']

def rep(path, out):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write('# %s -- selected metrics per captured launch (ncu --set full --clock-control none)\n' % path)
        for r in rows[2:]:
            f.write('\nkernel: %s\n' % r[hdr.index('Kernel Name')])
            for w in WANT:
                if w in hdr:
                    f.write('  %-75s %s %s\n' % (w, r[hdr.index(w)], units[hdr.index(w)]))

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    kn, val = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[val].replace(',', ''))
        except ValueError:
            continue
        name = r[kn].split('(')[0][-60:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    unit = rows[1][hdr.index('Metric Unit')]
    with open(out, 'w') as f:
        f.write('# %s -- per-kernel totals of gpu__time_duration.sum (%s); cold-cache, serialised: compare SHARES\n' % (path, unit))
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-62s launches %7d  total %14.1f  share %6.2f%%  avg %10.2f\n' % (name, n, t, 100 * t / tot, t / n))

if __name__ == '__main__':
    kind, src, dst = sys.argv[1:4]
    (rep if kind == 'rep' else launches)(src, dst)
