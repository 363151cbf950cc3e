"""Builds csrc/libftl.so for sm_100a with nvcc (in-tree, so the .so travels with the repo).

    python -m continiousenvironment_follower_leader_b200.build [--force]

Flags: -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false.  -fmad=false keeps float32 /
float64 operations separately rounded so results match the reference's numpy arithmetic; fused
multiply-adds are written explicitly (fma/fmaf) where they are wanted.  The per-NB step kernels are
separate translation units compiled in parallel.
"""
import concurrent.futures
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
OUT = os.path.join(CSRC, "libftl.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
HEADERS = ["ftl_device.cuh", "ftl_step.cuh", "ftl_rays.cuh", "ftl_state_io.cuh", "ftl_launch.h",
           os.path.join("..", "..", "include", "ftl.h")]
MAX_BEARS = 4


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _compile(args):
    src, obj, extra = args
    cmd = [NVCC] + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False, ptxas_info=False):
    hdr_time = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    extra_v = ["-Xptxas", "-v"] if ptxas_info else []
    for nb in range(MAX_BEARS + 1):
        jobs.append(("ftl_step_nb.cu", os.path.join(objdir, "ftl_step_nb%d.o" % nb), ["-DFTL_NB=%d" % nb] + extra_v))
    jobs.append(("ftl_capi.cu", os.path.join(objdir, "ftl_capi.o"), extra_v))
    jobs.append(("ftl_policy.cu", os.path.join(objdir, "ftl_policy.o"), extra_v))   # the rollout's fused policy kernel
    jobs.append(("ftl_policy_tc.cu", os.path.join(objdir, "ftl_policy_tc.o"), extra_v))   # ... on tcgen05 / tensor memory
    jobs.append(("ftl_scenario_gen.cpp", os.path.join(objdir, "ftl_scenario_gen.o"), []))   # host-only C++
    todo = [j for j in jobs if force or _mtime(j[1]) < max(hdr_time, _mtime(os.path.join(CSRC, j[0])))]
    logs = []
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for src, rc, log in ex.map(_compile, todo):
                logs.append(log)
                if verbose and log.strip():
                    print(log)
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
    objs = [j[1] for j in jobs]
    if todo or _mtime(OUT) < max(_mtime(o) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT, "\n".join(logs)


SMALL_LISTS_OUT = os.path.join(CSRC, "libftl_smalllists.so")
SMALL_LISTS_FLAGS = ["-DFTL_EDGE_CAP=24", "-DFTL_PAIR_CAP=40", "-DFTL_UNC_PER_ENV=2", "-DFTL_ALLOW_EDGE_OVERFLOW",
                     "-DFTL_SCAN_WIDE=1", "-DFTL_WALK_WIDE=1"]   # ... and one load in flight per lane in k_kin's warp scans / walks


def build_small_lists(force=False):
    """csrc/libftl_smalllists.so: the same sources with tiny shared lists in the ray kernel (24 edges, 40 pairs, 2 exact-pass
    records per env; the warp-cooperative scans and walks of k_kin in rounds of 32 points), a TEST build: every overflow path
    of the ray pass and every loop of the warp collectives runs on the GPU, where the lanes really are concurrent (tests/test_gpu_parity_gaps.py; the host build has the same lists in libftl_hostsim_smallcaps.so)."""
    build(force=False)   # the policy kernels and the scenario generator do not depend on the flags: their objects are shared
    hdr_time = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    objdir = os.path.join(CSRC, "build", "smalllists")
    os.makedirs(objdir, exist_ok=True)
    jobs = [("ftl_step_nb.cu", os.path.join(objdir, "nb%d.o" % nb), ["-DFTL_NB=%d" % nb] + SMALL_LISTS_FLAGS)
            for nb in range(MAX_BEARS + 1)]
    jobs.append(("ftl_capi.cu", os.path.join(objdir, "capi.o"), SMALL_LISTS_FLAGS))
    stamp = os.path.join(objdir, "flags.txt")   # objects made with other flags are stale
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(SMALL_LISTS_FLAGS):
        force = True
    todo = [j for j in jobs if force or _mtime(j[1]) < max(hdr_time, _mtime(os.path.join(CSRC, j[0])))]
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for src, rc, log in ex.map(_compile, todo):
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
        with open(stamp, "w") as f:
            f.write(" ".join(SMALL_LISTS_FLAGS))
    shared = [os.path.join(CSRC, "build", o) for o in ("ftl_policy.o", "ftl_policy_tc.o", "ftl_scenario_gen.o")]
    objs = [j[1] for j in jobs] + shared
    if todo or _mtime(SMALL_LISTS_OUT) < max(_mtime(o) for o in objs):
        r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", SMALL_LISTS_OUT] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return SMALL_LISTS_OUT


if __name__ == "__main__":
    out, log = build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas-info" in sys.argv)
    print("built", out)
    print("built", build_small_lists(force="--force" in sys.argv))
