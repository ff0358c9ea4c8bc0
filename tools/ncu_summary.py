"""Markdown table of the headline metrics of every kernel in an .ncu-rep (`ncu --set full`).
usage: python tools/ncu_summary.py REPORT.ncu-rep [name-filter]"""
import csv
import io
import subprocess
import sys

WANT = [
    ('us', 'gpu__time_duration.sum', 1.0),
    ('tensor pipe % (elapsed)', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 1.0),
    ('TC smem wavefronts %', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 1.0),
    ('issue active %', 'sm__inst_issued.avg.pct_of_peak_sustained_active', 1.0),
    ('LTS %', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
    ('DRAM %', 'FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
    ('DRAM read MB', 'dram__bytes_read.sum', None),
    ('DRAM write MB', 'dram__bytes_write.sum', None),
    ('L2 hit %', 'lts__t_sector_hit_rate.pct', 1.0),
    ('regs', 'launch__registers_per_thread', 1.0),
    ('dyn smem KB', 'launch__shared_mem_per_block_dynamic', None),
    ('stall long_sb', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 1.0),
    ('stall math', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 1.0),
    ('stall barrier', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 1.0),
    ('stall sleeping', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 1.0),
]


def main():
    rep = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')

    def col(metric):
        for i, h in enumerate(hdr):
            if h == metric:
                return i
        for i, h in enumerate(hdr):
            if h.endswith('.' + metric):
                return i
        return None

    cols = [(label, col(m), scale) for label, m, scale in WANT]
    print('| # | kernel | ' + ' | '.join(l for l, _, _ in cols) + ' |')
    print('|---|---|' + '---|' * len(cols))
    for n, r in enumerate(body):
        name = r[ki].replace('<unnamed>::', '').split('(')[0].replace('void ', '')
        if flt and flt not in name:
            continue
        vals = []
        for label, c, scale in cols:
            if c is None or c >= len(r) or r[c] in ('', 'no data'):
                vals.append('--')
                continue
            v = float(r[c].replace(',', ''))
            u = units[c]
            if scale is None:       # bytes of whatever unit -> MB / KB
                mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u.split('/')[0], 1)
                v = v * mult / (1e3 if 'KB' in label else 1e6)
            elif label == 'us':
                v = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'ns ': 1e-3}.get(u, 1.0) if u in ('ns', 'us', 'ms') else v
            vals.append(f'{v:.1f}')
        print(f'| {n} | {name} | ' + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    main()
