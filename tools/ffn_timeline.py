#!/usr/bin/env python
"""Timeline of ffn_tc_kernel from a -DAGH_FFN_TRACE build (CTA 0 of the first large launches).

    AGH_NVCC_EXTRA=-DAGH_FFN_TRACE python -c "from agh_b200 import _lib; _lib.build_library(force=True)"
    python tools/check_ffn.py --shapes 20x7 --time | python tools/ffn_timeline.py

Ops are numbered in MMA issue order, 16 per tile: A0 A1 B0 A2 B1 ... A7 B6 B7 (A_j: H chunk j = x W1_j^T; B_j: Y += H_j W2_j^T).
"""
import collections
import sys


def main():
    tr = collections.defaultdict(lambda: collections.defaultdict(dict))
    for ln in sys.stdin:
        if not ln.startswith('FTRACE '):
            continue
        _, l, role, idx, ns = ln.split()
        tr[int(l)][int(role)][int(idx)] = int(ns)
    names = ['A0', 'A1', 'B0', 'A2', 'B1', 'A3', 'B2', 'A4', 'B3', 'A5', 'B4', 'A6', 'B5', 'A7', 'B6', 'B7']
    for l in sorted(tr):
        r = tr[l]
        print('launch %d' % l)
        print('  op      start  +acc-free  +W-landed  +operand   +issue   | W slot free->TMA issue before op start')
        for i in range(32, 80):
            print('  %3d %-3s %7d %8d %9d %9d %9d   | %7d' % (i, names[i % 16], r[1][i], r[2][i] - r[1][i], r[3][i] - r[2][i], r[4][i] - r[3][i],
                                                         r[5][i] - r[4][i], r[0][i] - r[1][i]))
        per_tile = (r[5][111] - r[5][47]) / 4.0
        print('  tile period %.0f ns (ops 48..111)' % per_tile)
        print('  H epilogue per chunk: H ready -> TMEM read %d ns, wait for the Hs buffer %d ns, convert+store %d ns' % tuple(
            sum(r[a][c] - r[b][c] for c in range(16, 56)) / 40 for a, b in ((7, 6), (8, 7), (9, 8))))
        print('  Y epilogue per tile: x split done at %s; Y ready %s; TMEM read %s; stored %s' % tuple([r[k][t] for t in range(2, 6)] for k in (10, 11, 12, 13)))


if __name__ == '__main__':
    main()
