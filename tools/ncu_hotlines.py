"""Hot source lines of one kernel of an ncu report (stall samples, warp instructions, active lanes per instruction).

    python tools/ncu_hotlines.py gpurun_out/X.ncu-rep k_book [n_lines]
"""
import collections, csv, subprocess, sys

rep, kernel = sys.argv[1], sys.argv[2]
nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--print-source",
                      "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, per = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        continue
    if r[0].isdigit() and hdr:
        try:
            inst, s, ti = int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]]), int(r[hdr["Thread Instructions Executed"]])
        except Exception:
            continue
        p = per.setdefault((cur, int(r[0])), [0, 0, 0, r[1][:100]])
        p[0] += inst; p[1] += s; p[2] += ti
ts, ti = sum(v[1] for v in per.values()) or 1, sum(v[0] for v in per.values()) or 1
print("kernel %s: %d stall samples, %d warp instructions" % (kernel, ts, ti))
byfile = collections.Counter()
for (f, l), v in per.items():
    byfile[f] += v[1]
print("samples by file:", {k: "%.1f%%" % (100 * v / ts) for k, v in byfile.most_common()})
for (f, l), v in sorted(per.items(), key=lambda kv: -kv[1][1])[:nlines]:
    print("%-20s %4d  samples %5.1f%%  inst %5.1f%%  lanes %4.1f | %s" % (f, l, 100 * v[1] / ts, 100 * v[0] / ti, v[2] / max(v[0], 1), v[3]))
