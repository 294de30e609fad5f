#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: per kernel name the launch
count, mean and total duration.  Usage: launch_summary.py launches.csv [name-filter]"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    iN, iM, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    flt = sys.argv[2] if len(sys.argv) > 2 else ''
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= iV or r[iM] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r[iN]).replace('void ', '').replace('agh::', '').replace('(anonymous namespace)::', '')
        if flt and flt not in name:
            continue
        v = float(r[iV].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[iU], 1.0)
        a = agg.setdefault(name[:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-70s n=%4d mean %9.1f us total %10.1f us (%4.1f%%)' % (k, c, t / c, t, 100 * t / max(tot, 1e-9)))


if __name__ == '__main__':
    main()
