"""Key counters of an `ncu --set full` report as a markdown table (and DRAM bytes per launch as JSON).
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [label ...] > profiles/x.md"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__cluster_dim_x', 'cluster'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 % of peak'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM bytes'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM % of peak'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('sm__cycles_elapsed.max', 'cycles'),
]


def main():
    rep = sys.argv[1]
    labels = sys.argv[2:]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = rows[2:]
    names = [labels[k] if k < len(labels) else r[col['Kernel Name']].split('(')[0] for k, r in enumerate(launches)]
    print(f'# ncu --set full --clock-control none: `{rep}`\n')
    print('| metric | ' + ' | '.join(names) + ' |')
    print('|---|' + '---|' * len(names))
    traffic = {}
    for key, title in KEYS:
        if key not in col:
            continue
        i = col[key]
        print(f'| {title} ({units[i]}) | ' + ' | '.join(r[i] for r in launches) + ' |')
    for k, r in enumerate(launches):
        def val(key):
            i = col[key]
            v = float(r[i].replace(',', ''))
            u = units[i].lower()
            return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
        traffic[names[k]] = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
    print('\nDRAM bytes per launch (read + write): `' + json.dumps(traffic) + '`')


if __name__ == '__main__':
    main()
