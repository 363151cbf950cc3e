"""Time the two kernels for several builds of libftl.so (FTL_LIB override); diagnostic."""
import glob, os, subprocess, sys
libs = [None] + sorted(glob.glob("tools/libftl_*.so"))
for lib in libs:
    env = dict(os.environ)
    if lib: env["FTL_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.'); import sweep_f; a,b=sweep_f.run(65536,10,ref_pool=True); print('k_step %.4f k_rays %.4f'%(a,b))"], env=env, capture_output=True, text=True)
    print(lib or "default", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
