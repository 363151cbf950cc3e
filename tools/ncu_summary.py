"""Summarise an `ncu --set full` report for profiles/: one block of metrics per kernel plus DRAM traffic per launch.

usage: python tools/ncu_summary.py report.ncu-rep profiles/rNN_ncu_summary.txt profiles/rNN_traffic.json
"""
import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def to_bytes(v, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * scale.get(unit, 1.0)


def main():
    rep, out_txt, out_json = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    traffic, lines = {"warp_inst": {}, "lanes_per_inst": {}}, [sys.argv[4] if len(sys.argv) > 4 else "ncu --set full --clock-control none, bench.py cfg3 workload"]
    seen = set()
    for row in rows[2:]:
        d = dict(zip(head, row))
        u = dict(zip(head, units))
        name = d["Kernel Name"]
        short = name.split("(")[0].replace("void ", "")
        if short in seen:
            continue
        seen.add(short)
        lines.append("---- " + name)
        for m in METRICS:
            if m in d and d[m] != "":
                lines.append("  %-86s %s %s" % (m, d[m], u.get(m, "")))
        key = {"k_kin": "k_step"}.get(short.split("<")[0], short.split("<")[0])   # bench.py's kernel groups
        traffic[key] = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + \
            to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        traffic["warp_inst"][key] = float(d["smsp__inst_executed.sum"].replace(",", ""))
        traffic["lanes_per_inst"][key] = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"].replace(",", ""))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_json, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic))


main()
