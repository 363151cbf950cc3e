"""Stall samples and warp instructions of one kernel of an ncu report aggregated by source FUNCTION (line ranges of the
FTL_HD functions of csrc/*.cuh).   python tools/ncu_funcs2.py X.ncu-rep k_kin"""
import bisect, collections, csv, os, re, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "continiousenvironment_follower_leader_b200", "csrc")
funcs = {}
for f in os.listdir(CSRC):
    if not f.endswith((".cuh", ".cu")): continue
    L = []
    for i, l in enumerate(open(os.path.join(CSRC, f)), 1):
        m = re.match(r'^(?:FTL_HD|FTL_HD_NOINLINE|static|__global__|__device__)[\w\s:<>\*&,]*?\b(\w+)\(', l)
        if m: L.append((i, m.group(1)))
    funcs[f] = L
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
cur, hdr = None, None
S, I = collections.Counter(), collections.Counter()
seen_kernel = 0
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {k: i for i, k in enumerate(r)}; continue
    if r[0].isdigit() and hdr:
        try: inst, s = int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]])
        except Exception: continue
        L = funcs.get(cur)
        name = cur
        if L:
            k = bisect.bisect_right([x[0] for x in L], int(r[0])) - 1
            if k >= 0: name = cur.replace("ftl_", "").split(".")[0] + ":" + L[k][1]
        S[name] += s; I[name] += inst
ts, ti = sum(S.values()) or 1, sum(I.values()) or 1
print("%s: %d samples, %d warp instructions" % (kernel, ts, ti))
for k, v in S.most_common(45):
    print("%-44s samples %5.1f%%   inst %5.1f%%" % (k, 100 * v / ts, 100 * I[k] / ti))
