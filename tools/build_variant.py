"""Build tools/libftl_<tag>.so with extra -D flags (A/B timing with tools/ab_libs.py; AB_LIB selects a build for tools/sweep_f.py).

    python tools/build_variant.py <tag> [-DNAME=VALUE ...] [--nb-only]
"""
import concurrent.futures, os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from continiousenvironment_follower_leader_b200 import build as B

def main():
    tag, flags = sys.argv[1], [a for a in sys.argv[2:] if a.startswith("-")]
    objdir = os.path.join(B.CSRC, "build", "var_" + tag)
    os.makedirs(objdir, exist_ok=True)
    jobs = [("ftl_step_nb.cu", os.path.join(objdir, "nb%d.o" % nb), ["-DFTL_NB=%d" % nb] + flags) for nb in range(B.MAX_BEARS + 1)]
    jobs.append(("ftl_capi.cu", os.path.join(objdir, "capi.o"), flags))
    jobs.append(("ftl_policy.cu", os.path.join(objdir, "policy.o"), flags))
    jobs.append(("ftl_policy_tc.cu", os.path.join(objdir, "policy_tc.o"), flags))
    jobs.append(("ftl_scenario_gen.cpp", os.path.join(objdir, "gen.o"), []))
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for src, rc, log in ex.map(B._compile, jobs):
            if rc != 0: raise SystemExit("nvcc failed on %s:\n%s" % (src, log))
            if "-Xptxas" in flags and "nb1" in log: print(log)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libftl_%s.so" % tag)
    r = subprocess.run([B.NVCC] + B.ARCH + ["-shared", "-o", out] + [j[1] for j in jobs], capture_output=True, text=True)
    if r.returncode != 0: raise SystemExit(r.stdout + r.stderr)
    print("built", out)

if __name__ == "__main__":
    main()
