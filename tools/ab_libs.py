"""Time the two kernels for several builds of libftl.so (the tool passes the variant to FtlBatchEnv(lib_path=...) through AB_LIB) on one box, with a checksum of the
state after the run so that a variant that changes results is visible; diagnostic.

    python tools/ab_libs.py [glob]         (default glob: tools/libftl_*.so)
"""
import glob, os, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "tools/libftl_*.so"
libs = [None] + sorted(glob.glob(pat))
CODE = r'''
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import sweep_f, torch, hashlib
a,b,h = sweep_f.run(65536,10,ref_pool=True,checksum=True)
k=sweep_f.run.last_kernels
print('k_kin %.4f k_book %.4f k_rays %.4f sum %.4f  step(no events) %.4f  %s' % (k['k_kin'],k['k_book'],k['k_rays'],a+b,sweep_f.run.last_total_ms,h))
'''
for rep in range(int(os.environ.get("AB_REPS", "1"))):
    for lib in libs:
        env = dict(os.environ)
        if lib: env["AB_LIB"] = os.path.abspath(lib)
        out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
        print("%-28s" % (os.path.basename(lib) if lib else "default"), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-600:], flush=True)
