"""Attribute an ncu source-page CSV (SASS view) to CUDA source lines using nvdisasm line info.

usage: python tools/ncu_lines.py k.csv file.cubin kernel_mangled_substring [top]
Aggregates Instructions Executed / samples per (file, line) of the innermost inlined location.
"""
import csv, re, subprocess, sys, collections

def main():
    rows = list(csv.reader(open(sys.argv[1])))
    cubin, kname = sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    h = rows[1]; ix = {k: i for i, k in enumerate(h)}
    R = [r for r in rows[2:] if len(r) == len(h)]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    loc, cur, inside, offs = {}, None, False, {}
    sect = None
    for ln in dis:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            sect = m.group(1); continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and sect:
            offs.setdefault(sect, []).append((int(m.group(1), 16), cur, m.group(2).strip()))
    main_sect = [s for s in offs if kname in s][0]
    base = int(R[0][ix["Address"]], 16)
    table = {o: (l, t) for o, l, t in offs[main_sect]}
    def col(r, k):
        try: return float(r[ix[k]])
        except ValueError: return 0.0
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    tot_i = sum(col(r, "Instructions Executed") for r in R); tot_s = sum(col(r, "# Samples") for r in R)
    for r in R:
        off = int(r[ix["Address"]], 16) - base
        l = table.get(off, (("callee", 0), ""))[0] if off in table else ("callee/other", 0)
        a = agg[l]
        a[0] += col(r, "Instructions Executed"); a[1] += col(r, "# Samples"); a[2] += col(r, "Thread Instructions Executed"); a[3] += 1
    print("total warp-instr %d samples %d" % (tot_i, tot_s))
    for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s:%4d  instr %5.2f%%  samples %5.2f%%  lanes %4.1f  sass %3d" % (l[0], l[1], 100 * a[0] / tot_i, 100 * a[1] / tot_s, a[2] / max(a[0], 1), a[3]))

main()
