#!/usr/bin/env python
"""Timeline of mha_tc_kernel from a -DAGH_MHA_TRACE build (CTA 0, first instances of the first large launches).

    AGH_NVCC_EXTRA=-DAGH_MHA_TRACE python -c "from agh_b200 import _lib; _lib.build_library(force=True)"
    python tools/check_ffn.py --mha --shapes 20x1 --time | python tools/mha_timeline.py
"""
import collections
import sys


def main():
    tr = collections.defaultdict(lambda: collections.defaultdict(dict))
    for ln in sys.stdin:
        if not ln.startswith('ATRACE '):
            continue
        _, l, role, idx, ns = ln.split()
        tr[int(l)][int(role)][int(idx)] = int(ns)
    for l in sorted(tr):
        r = tr[l]
        print('launch %d' % l)
        print('  loads j (3 per instance: Q K V): buffer free -> raw landed -> split done')
        for j in range(0, 15):
            print('   j=%2d %s  free %7d  landed +%5d  split +%5d' % (j, 'QKV'[j % 3], r[0][j], r[1][j] - r[0][j], r[2][j] - r[1][j]))
        print('  head G: S start, +waits, +issue | PV(G) start, +waits, +issue | softmax(G): loop top, S seen, P written')
        for G in range(8, 40):
            gi, c = G & 1, G >> 1
            print('   G=%2d  S %7d +%5d +%4d | PV %7d +%5d +%4d | sm %7d  %7d  %7d (softmax %5d)' % (
                G, r[3][G], r[4][G] - r[3][G], r[5][G] - r[4][G], r[6][G], r[7][G] - r[6][G], r[8][G] - r[7][G],
                r[9 + 3 * gi][c], r[10 + 3 * gi][c], r[11 + 3 * gi][c], r[11 + 3 * gi][c] - r[10 + 3 * gi][c]))
        if 16 in r:
            print('  inside the softmax of group 0 (warp of lane quarter 0), ns after the scores were seen:')
            print('   head: pass-1 loads landed | max done | chunk 0 / 1 / 2 loads landed | all chunks done | stores done | P published')
            for c in range(8, 20):
                t0 = r[10][c]
                print('   %4d: ' % (2 * c) + ' '.join('%6d' % (r[k][c] - t0) for k in (16, 17, 18, 19, 20, 21, 22, 11)))
        print('  per instance: %.0f ns (heads 16..48)' % ((r[5][48] - r[5][16]) / 4.0))


if __name__ == '__main__':
    main()
