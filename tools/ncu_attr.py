"""Attribute an ncu source-page CSV to device functions, sub-functions and source lines, with the average number of
active lanes per warp instruction (finds work that a few lanes do while the rest of the warp waits).

    ncu -i X.ncu-rep --page source --csv --kernel-name k_step > k.csv
    nvcc ... -cubin csrc/ftl_step_nb.cu -DFTL_NB=1 -o nb1.cubin        (the same sources the report was captured from)
    python tools/ncu_attr.py k.csv nb1.cubin k_stepILi1 [n_lines]
"""
import collections, csv, os, re, subprocess, sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "continiousenvironment_follower_leader_b200", "csrc")


def ftab(path):
    t = []
    for n, l in enumerate(open(path), 1):
        if re.match(r"^(FTL_HD|FTL_HD_NOINLINE|template|static|__device__|__global__|inline)\b", l):
            m = re.search(r"(\w+)\s*\(", l)
            if m and m.group(1) not in ("defined", "__launch_bounds__"):
                t.append((n, m.group(1)))
        elif re.match(r"^k_\w+\(", l):
            t.append((n, l.split("(")[0]))
    return t


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    cubin, kname = sys.argv[2], sys.argv[3]
    nlines = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    FT = {f: ftab(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".cu"))}

    def fn(file, line):
        name = "?"
        for s, n in FT.get(file, []):
            if line >= s:
                name = n
        return name if file in FT else file

    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    R = [r for r in rows[2:] if len(r) == len(h)]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    sect, sub, cur, tab = None, "main", None, []
    for ln in dis:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            sect, sub = m.group(1), "main"
            continue
        if sect is None or kname not in sect:
            continue
        m = re.match(r"^\$.*\$(\S+):", ln) or re.match(r"^(\$?__\w+):", ln)
        if m:   # a device function the kernel calls (not inlined)
            nm = m.group(1)
            mm = re.search(r"ftl(\d+)(\w+)", nm)
            sub = (mm.group(2)[:int(mm.group(1))] if mm else nm)[:40]
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            tab.append((sub, cur))
    R = R[:len(tab)]
    if len(R) != len(tab):
        raise SystemExit("the cubin has %d instructions, the report %d: not the same build" % (len(tab), len(R)))

    def col(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0
    toti = sum(col(r, "Instructions Executed") for r in R)
    tots = sum(col(r, "# Samples") for r in R)
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    lines = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for r, (sub, cur) in zip(R, tab):
        a = agg[(sub, fn(*cur) if cur else "?")]
        a[0] += col(r, "Instructions Executed"); a[1] += col(r, "# Samples"); a[2] += col(r, "Thread Instructions Executed"); a[3] += 1
        if cur:
            l = lines[(sub,) + cur]
            l[0] += col(r, "Instructions Executed"); l[1] += col(r, "# Samples"); l[2] += col(r, "Thread Instructions Executed")
    print("warp instructions %.0f, samples %.0f" % (toti, tots))
    print("%-22s %-26s %7s %8s %6s %6s" % ("called function", "source function", "instr%", "samples%", "lanes", "SASS"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-22s %-26s %6.2f%% %7.2f%% %6.1f %6d" % (k[0][:22], k[1][:26], 100 * a[0] / toti, 100 * a[1] / tots, a[2] / max(a[0], 1), a[3]))
    print("--- source lines")
    for k, l in sorted(lines.items(), key=lambda kv: -kv[1][1])[:nlines]:
        print("%-22s %-18s:%-5d instr %5.2f%% samples %5.2f%% lanes %4.1f" % (k[0][:22], k[1], k[2], 100 * l[0] / toti, 100 * l[1] / tots, l[2] / max(l[0], 1)))


if __name__ == "__main__":
    main()
