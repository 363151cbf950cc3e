"""Diagnostic: how often, and how clustered, the exact green-zone scans are in the bench workload (cfg3, reference pool,
uniform random actions, auto-reset).  Runs the host build of the device functions with an event hook on FTL_COUNT
(tools/scan_stats/scan_stats.cpp); prints scans per env-frame, per env-step and per group of 32 envs ("warp") per step.

    python tools/scan_stats.py [n_envs] [steps] [settle]
"""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from continiousenvironment_follower_leader_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
settle = int(sys.argv[3]) if len(sys.argv) > 3 else 150
src = os.path.join(ROOT, "tools", "scan_stats")
so = "/tmp/libscanstats.so"
subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-DFTL_COUNT_HOOK",
                       "-Wno-unknown-pragmas", "-x", "c++", "-shared", "-o", so, os.path.join(src, "scan_stats.cpp"),
                       os.path.join(ROOT, "continiousenvironment_follower_leader_b200", "csrc", "ftl_scenario_gen.cpp"),
                       "-lm", "-lpthread"])
L = ctypes.CDLL(so)
L.scan_events.restype = ctypes.c_longlong
L.scan_events.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
lib = capi.bind(L, [k for k in capi.SIGNATURES if hasattr(L, k)])
gc = bench.workload_config()
pool, kind = bench.workload_pool(gc)
env = capi.HostEnv(gc, n, lib=lib)
env.upload_scenarios(pool)
env.reset()
rng = np.random.default_rng(0)
lo, hi = gc.action_bounds()
F = gc.c.frames_per_step


def drain():
    cnt = L.scan_events(None, 0)
    buf = np.zeros(cnt, np.int64)
    L.scan_events(buf.ctypes.data, cnt)
    return buf >> 32, buf & 0xffffffff


for t in range(settle):
    env.step(rng.uniform(lo, hi, size=(n, 2)).astype(np.float32))
drain()
per_frame_g, per_frame_a, pts = [], [], []
walks = 0
for t in range(steps):
    env.step(rng.uniform(lo, hi, size=(n, 2)).astype(np.float32))
    k, v = drain()
    # events: 3 = one green_flags call (env-frame); 0 = a scan (followed by 1 = its points); 5 = whole-trail scan
    idx = np.cumsum(k == 3) - 1                       # env-frame index of every event (resets add calls: clip below)
    nfr = int((k == 3).sum())
    walks += int((k == 2).sum())
    g = np.bincount(idx[k == 0], minlength=nfr)       # all scans (green + whole-trail)
    a = np.bincount(idx[k == 5], minlength=nfr)
    p = np.bincount(idx[k == 1], weights=v[k == 1], minlength=nfr)
    per_frame_g.append(g[:n * F].reshape(n, F)); per_frame_a.append(a[:n * F].reshape(n, F)); pts.append(p[:n * F].reshape(n, F))
G = np.stack(per_frame_g).astype(np.int64)   # [steps, n, F] scans per env-frame (0, 1 or 2)
A = np.stack(per_frame_a)
P = np.stack(pts)
print("pool %s, %d envs, %d steps after %d settle steps" % (kind, n, steps, settle))
print("env-frames with >= 1 scan: %.2f %% (green-window scans %.2f %%, whole-trail %.2f %%), points per scan %.0f"
      % (100 * (G > 0).mean(), 100 * ((G - A) > 0).mean(), 100 * (A > 0).mean(), P.sum() / max(G.sum(), 1)))
es = G.sum(axis=2)    # scans per env-step
print("scans per env-step: mean %.3f; histogram 0..10+: %s" % (es.mean(), np.bincount(np.minimum(es.ravel(), 10), minlength=11) / es.size))
w = G.reshape(steps, n // 32, 32, F)
wf = w.sum(axis=2)            # scans per warp-frame
ws = wf.sum(axis=2)           # scans per warp-step
print("scans per warp-step: mean %.2f, median %.0f, p90 %.0f, p99 %.0f, max %d" % (ws.mean(), np.median(ws), np.percentile(ws, 90),
                                                                                  np.percentile(ws, 99), ws.max()))
pw = P.reshape(steps, n // 32, 32, F).sum(axis=(2, 3))
print("points scanned per warp-step: mean %.0f, p90 %.0f, p99 %.0f, max %d" % (pw.mean(), np.percentile(pw, 90), np.percentile(pw, 99), pw.max()))
print("exact float32 window walks (green_lo_exact): %.4f per env-step" % (walks / (steps * n)))
# envs that scan in k of the 10 frames of a step
print("share of all scans made by envs with >= 3 scans in the step: %.1f %%" % (100 * es[es >= 3].sum() / max(es.sum(), 1)))
