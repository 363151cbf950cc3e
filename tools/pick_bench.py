import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"]/1e6,2), d["ms_per_step"], d["kernels_ms_per_step"], d["clocks"]["sm_mhz"], round(d["e2e"]["value"]/1e6,2))
