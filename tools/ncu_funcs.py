"""Attribute an ncu source-page CSV (SASS view) to the functions of csrc/*.cuh via nvdisasm line info.

usage: python tools/ncu_funcs.py k.csv file.cubin kernel_substring [top]
"""
import collections, csv, os, re, subprocess, sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "continiousenvironment_follower_leader_b200", "csrc")


def ftab(path):
    t = []
    for n, l in enumerate(open(path), 1):
        if re.match(r"^(FTL_HD|FTL_HD_NOINLINE|template|static|__device__|__global__|inline)\b", l):
            m = re.search(r"(\w+)\s*\(", l)
            if m and m.group(1) not in ("defined",):
                t.append((n, m.group(1)))
    return t


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    cubin, kname = sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    h = rows[1]; ix = {k: i for i, k in enumerate(h)}
    R = [r for r in rows[2:] if len(r) == len(h)]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    sect, cur, offs = None, None, {}
    for ln in dis:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: sect = m.group(1); continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and sect: offs.setdefault(sect, []).append((int(m.group(1), 16), cur, m.group(2).strip()))
    ms = [s for s in offs if kname in s][0]
    tab = {o: (l, t) for o, l, t in offs[ms]}
    base = int(R[0][ix["Address"]], 16)
    bad = sum(1 for r in R[:len(offs[ms])] if (int(r[ix["Address"]], 16) - base) not in tab or
              r[1].strip().split()[-1][:6] != tab[int(r[ix["Address"]], 16) - base][1].split()[-1][:6])
    print("rows %d, sass in cubin %d, mismatching %d" % (len(R), len(offs[ms]), bad))
    FT = {f: ftab(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".cu"))}

    def fn(file, line):
        if file not in FT: return file
        name = "?"
        for s, n in FT[file]:
            if line >= s: name = n
        return name

    def col(r, k):
        try: return float(r[ix[k]])
        except ValueError: return 0.0
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0, 0.0, 0.0])
    toti = sum(col(r, "Instructions Executed") for r in R); tots = sum(col(r, "# Samples") for r in R)
    for r in R:
        off = int(r[ix["Address"]], 16) - base
        t = tab.get(off)
        key = fn(*t[0]) if t and t[0] else "callee/other"
        a = agg[key]
        a[0] += col(r, "Instructions Executed"); a[1] += col(r, "# Samples"); a[2] += col(r, "Thread Instructions Executed")
        a[3] += 1 if col(r, "Instructions Executed") > 0 else 0
        a[4] += col(r, "stall_no_inst"); a[5] += col(r, "stall_long_sb")
    print("%-30s %7s %8s %6s %9s %8s %8s" % ("function", "instr%", "samples%", "lanes", "live sass", "no_inst%", "long_sb%"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-30s %6.1f%% %7.1f%% %6.1f %9d %7.1f%% %7.1f%%" % (k, 100 * a[0] / toti, 100 * a[1] / tots, a[2] / max(a[0], 1), a[3],
                                                             100 * a[4] / tots, 100 * a[5] / tots))


main()
