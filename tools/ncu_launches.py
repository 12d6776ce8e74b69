#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
shares, and optionally the first N launches in order.  usage: ncu_launches.py file.csv [N]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hi]
    ki, vi, ui, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('Grid Size')
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(',', ''))
        t = t / 1e3 if r[ui] == 'ns' else (t * 1e3 if r[ui] == 'ms' else t)
        seq.append((r[ki].split('(')[0], t, r[gi]))
    return seq


if __name__ == '__main__':
    seq = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, t, _ in seq:
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    print('total %.1f us over %d launches' % (tot, len(seq)))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-44s %6d launches %10.1f us %5.1f%%' % (k[:44], v[0], v[1], 100 * v[1] / tot))
    if len(sys.argv) > 2:
        for s in seq[:int(sys.argv[2])]:
            print('%-32s %9.2f us  grid %s' % s)
