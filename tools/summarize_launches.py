"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into one training step (k_gather .. k_adam).
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import csv
import sys
from collections import defaultdict


def main(path):
    lines = [ln for ln in open(path) if not ln.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ni = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    gi = hdr.index('Grid Size')
    bi = hdr.index('Block Size')
    seq = []
    for row in r:
        if len(row) > vi and row[ni] == 'gpu__time_duration.sum':
            name = row[ki].replace('void ', '').split('(')[0].replace('jb::', '')
            seq.append((name, float(row[vi].replace(',', '')) / 1000.0, row[gi], row[bi]))
    starts = [i for i, s in enumerate(seq) if 'k_gather' in s[0] or 'k_begin' in s[0]]
    if len(starts) < 2:
        print('no complete step found')
        return
    s, e = starts[-2], starts[-1]   # a late, warm step
    step = seq[s:e]
    tot = sum(x[1] for x in step)
    print(f'# One optimizer step, kernel by kernel ({path})\n')
    print('`ncu --metrics gpu__time_duration.sum --clock-control none` on `bench.py --steps 6 --warmup 8`: per-launch times are')
    print('cold-cache and serialised (the graph replays them back to back), so compare SHARES, not absolutes.\n')
    print(f'{len(step)} launches, {tot:.1f} us serialised.\n')
    print('| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|')
    for i, (n, v, g, b) in enumerate(step):
        print(f'| {i} | `{n}` | {g} | {b} | {v:.2f} | {100 * v / tot:.1f}% |')
    agg = defaultdict(lambda: [0.0, 0])
    for n, v, _, _ in step:
        agg[n][0] += v
        agg[n][1] += 1
    print('\n| kernel | launches | us | share |\n|---|---|---|---|')
    for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f'| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |')


if __name__ == '__main__':
    main(sys.argv[1])
