#!/usr/bin/env python
"""Per-stage timeline of gemm_tc_kernel from a -DAGH_TC_TRACE build.

    AGH_NVCC_EXTRA=-DAGH_TC_TRACE python -c "from agh_b200 import _lib; _lib.build_library(force=True)"
    python tools/bench_encoder.py --iters 1 | python tools/gemm_timeline.py

Reads the `TRACE launch EPI WRES role idx ns` lines CTA (0,0) prints and reports, per traced launch, the mean time
(ns) between pipeline hand-offs in steady state (k-blocks 8.. of the trace):
  role 0 producer: stage free -> TMA issued      role 1 splitter: raw landed      role 2 splitter: split done
  role 3 MMA: operands ready                      role 4 MMA: issued + committed
  role 5 epilogue: accumulator ready              role 6 epilogue: TMEM read out   role 7 epilogue: tile stored
"""
import collections
import sys

EPI = {0: 'plain (QKV)', 1: 'bias+ReLU (FF1)', 2: 'residual+BN', 3: 'decoder projection'}


def main():
    tr = collections.defaultdict(lambda: collections.defaultdict(dict))
    meta = {}
    for ln in sys.stdin:
        if not ln.startswith('TRACE '):
            continue
        _, l, epi, wres, role, idx, ns = ln.split()
        tr[int(l)][int(role)][int(idx)] = int(ns)
        meta[int(l)] = (int(epi), int(wres))
    for l in sorted(tr):
        r = tr[l]
        epi, wres = meta[l]
        # k-blocks per tile and ring depth: FP16 build (default) K = 64 per stage, 4 / 3 stages; --tf32: K = 32, 3 stages
        tf32 = '--tf32' in sys.argv
        nkb = (4 if wres else 16) if tf32 else (2 if wres else 8)
        depth = 3 if (tf32 or not wres) else 4
        kst = 32 if tf32 else 64
        n = min(len(r[0]), len(r[1]), len(r[2]), len(r[3]), len(r[4]))
        ks = [i for i in range(8, n)]
        if not ks:
            continue

        def mean(f):
            v = [f(i) for i in ks]
            return sum(v) / len(v)

        per_kb = (r[4][ks[-1]] - r[4][ks[0]]) / (len(ks) - 1)
        print('launch %d  %s  K=%d: %.0f ns per k-block = %.2f us per tile in steady state' % (l, EPI[epi], kst * nkb, per_kb, per_kb * nkb / 1e3))
        print('   TMA issue -> raw landed        %6.0f ns' % mean(lambda i: r[1][i] - r[0][i]))
        print('   raw landed -> split done       %6.0f ns' % mean(lambda i: r[2][i] - r[1][i]))
        print('   split done -> MMA sees it      %6.0f ns' % mean(lambda i: r[3][i] - r[2][i]))
        print('   MMA issue + commit             %6.0f ns' % mean(lambda i: r[4][i] - r[3][i]))
        print('   commit -> stage refilled (TMA issue of k-block i+depth)  %6.0f ns' % mean(lambda i: r[0][i + depth] - r[4][i] if i + depth in r[0] else 0))
        tiles = [t for t in range(2, min(len(r[5]), len(r[6]), len(r[7])))]
        if tiles:
            def tmean(f):
                v = [f(t) for t in tiles]
                return sum(v) / len(v)
            print('   epilogue: last MMA commit -> accumulator ready  %6.0f ns' % tmean(lambda t: r[5][t] - r[4][min(t * nkb + nkb - 1, n - 1)]))
            print('   epilogue: accumulator ready -> TMEM read out    %6.0f ns' % tmean(lambda t: r[6][t] - r[5][t]))
            print('   epilogue: TMEM read out -> tile stored          %6.0f ns' % tmean(lambda t: r[7][t] - r[6][t]))
            print('   epilogue: tile period                            %6.0f ns' % ((r[7][tiles[-1]] - r[7][tiles[0]]) / max(len(tiles) - 1, 1)))


if __name__ == '__main__':
    main()
