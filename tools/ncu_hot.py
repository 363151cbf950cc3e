"""Summarise an ncu source-page CSV: stall shares by execution-count bucket and the top stalled SASS lines.

usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > k.csv ; python tools/ncu_hot.py k.csv [warps]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    warps = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    R = [r for r in rows[2:] if len(r) == len(h)]

    def col(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0

    tot = sum(col(r, '# Samples') for r in R)
    toti = sum(col(r, 'Instructions Executed') for r in R)
    print('SASS rows %d, samples %d, warp-instructions %d' % (len(R), tot, toti))
    for key in ['stall_long_sb', 'stall_no_inst', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_selected']:
        print('  %-24s %5.1f%%' % (key, 100 * sum(col(r, key) for r in R) / tot))
    for lo, hi in [(0, 1), (1, warps + 1), (warps + 1, 2 * warps + 100), (2 * warps + 100, 5 * warps), (5 * warps, 10.5 * warps),
                   (10.5 * warps, 22 * warps), (22 * warps, 35 * warps), (35 * warps, 1e12)]:
        sel = [r for r in R if lo <= col(r, 'Instructions Executed') < hi]
        print('ex/warp in [%.1f,%.1f): n=%5d samples=%5.1f%% instr=%5.1f%% long_sb=%5d no_inst=%5d wait=%5d' % (
            lo / warps, hi / warps, len(sel), 100 * sum(col(r, '# Samples') for r in sel) / tot,
            100 * sum(col(r, 'Instructions Executed') for r in sel) / toti, sum(col(r, 'stall_long_sb') for r in sel),
            sum(col(r, 'stall_no_inst') for r in sel), sum(col(r, 'stall_wait') for r in sel)))
    top = sorted(range(len(R)), key=lambda i: -col(R[i], '# Samples'))[:40]
    for i in top:
        r = R[i]
        print('%5d s=%4.0f lsb=%4.0f ni=%3.0f ex/w=%6.2f  %-56s | prev: %s' % (
            i, col(r, '# Samples'), col(r, 'stall_long_sb'), col(r, 'stall_no_inst'), col(r, 'Instructions Executed') / warps,
            r[1].strip()[:56], R[i - 1][1].strip()[:44]))


if __name__ == '__main__':
    main()
