"""Per-phase instruction budget of k_rays from an ncu report's source page (warp instructions per env, active lanes).

    python tools/ncu_rays_budget.py gpurun_out/X.ncu-rep [n_envs] > profiles/rNN_rays_budget.txt

Phases are found in the CURRENT csrc/ftl_rays.cuh by function and by the `// ----` markers inside rays_warp / ray_flush, so
the report must have been captured from the same sources.
"""
import collections, csv, os, re, subprocess, sys

rep = sys.argv[1]
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "continiousenvironment_follower_leader_b200", "csrc", "ftl_rays.cuh")
lines = open(SRC).read().splitlines()

# (first line, label) in file order; a line belongs to the last entry at or before it
marks = []
for n, l in enumerate(lines, 1):
    m = re.match(r"^(?:FTL_HD|FTL_HD_NOINLINE|inline)\s+\S.*?\b(\w+)\s*\(", l)
    if m:
        marks.append((n, "fn:" + m.group(1)))
    elif "// ---- A2: edges" in l:
        marks.append((n, "A2 edges -> pairs"))
    elif "// ---- B: uniform pair tests" in l:
        marks.append((n, "B pair loop"))
    elif "// ---- setup:" in l:
        marks.append((n, "setup"))
    elif "// ---- A1: rectangles" in l:
        marks.append((n, "A1 rectangle rounds"))
    elif "// ---- A1: corridor sides" in l:
        marks.append((n, "A1 corridor loop + caps"))
    elif "// ---- out:" in l:
        marks.append((n, "out"))
LABEL = {"fn:diff_of_products": "B seg_hit", "fn:seg_hit": "B seg_hit", "fn:atan2_deg_approx": "A2 atan2",
         "fn:unc_push": "B merge/unc", "fn:hit_merge": "B merge/unc", "fn:edge_ray_test": "B merge/unc",
         "fn:edge_inline": "A1 overflow (edge_inline)", "fn:edge_append": "A1 edge_append", "fn:rect_append": "A1 rect_append",
         "fn:seg_append": "A1 seg_append", "fn:ray_flush": "A2 edges -> pairs", "fn:rays_warp": "setup",
         "fn:ray_rows_write_fused": "out", "fn:smem_atomic_add": "shared atomics", "fn:smem_atomic_min": "shared atomics",
         "fn:f2i_bits": "B merge/unc", "fn:i2f_bits": "out"}


def phase(f, l):
    if f != "ftl_rays.cuh":
        return {"device_atomic_functions.hpp": "shared atomics", "ftl_device.cuh": "setup (float64 sincos)",
                "ftl_capi.cu": "kernel prologue / flag wait", "sm_30_intrinsics.hpp": "warp sync"}.get(f, f)
    lab = "other"
    for n, name in marks:
        if n <= l:
            lab = name
        else:
            break
    return LABEL.get(lab, lab)


out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_rays", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, ph, pht = None, None, collections.Counter(), collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
    elif r[0].isdigit() and hdr:
        try:
            inst, ti = int(r[hdr["Instructions Executed"]]), int(r[hdr["Thread Instructions Executed"]])
        except Exception:
            continue
        k = phase(cur, int(r[0]))
        ph[k] += inst
        pht[k] += ti
tot = sum(ph.values())
print("k_rays instruction budget, %s: %.0f warp instructions per env (%d envs), %.1f active lanes per instruction"
      % (os.path.basename(rep), tot / n_envs, n_envs, sum(pht.values()) / tot))
print("%-34s %12s %7s %7s" % ("phase", "inst / env", "share", "lanes"))
for k, v in ph.most_common():
    print("%-34s %12.1f %6.1f%% %7.1f" % (k, v / n_envs, 100.0 * v / tot, pht[k] / max(v, 1)))
