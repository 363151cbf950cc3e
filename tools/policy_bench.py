"""Device time of ftl_policy_mlp (65 536 x 240 -> 128 -> 128 -> 3) for every tools/libftl_pol_*.so variant, the default
library's tcgen05 kernel and its mma.sync kernel (FTL_POLICY_IMPL=mma); CUDA events around 200 back-to-back launches."""
import ctypes as C, glob, os, subprocess, sys
CODE = r'''
import ctypes as C, os, sys
import numpy as np, torch
torch.manual_seed(0)
sys.path.insert(0, ".")
from continiousenvironment_follower_leader_b200 import abi, capi
from continiousenvironment_follower_leader_b200.rollout import MlpPolicy
L = capi.load(os.environ.get("POL_LIB") or None)
n, D, A = 65536, 240, 2
pol = MlpPolicy(D, -np.ones(A, np.float32), np.ones(A, np.float32), seed=1).cuda()
obs = torch.rand(4, n, D, device="cuda"); noise = torch.randn(n, A, device="cuda")
keep = {"w1": pol.body[0].weight.detach().to(torch.bfloat16).contiguous(), "b1": pol.body[0].bias.detach().contiguous(),
        "w2": pol.body[2].weight.detach().to(torch.bfloat16).contiguous(), "b2": pol.body[2].bias.detach().contiguous(),
        "w3": pol.head.weight.detach().to(torch.bfloat16).contiguous(), "b3": pol.head.bias.detach().contiguous(),
        "ns": pol.log_std.detach().exp().contiguous(), "mid": pol.act_mid.contiguous(), "half": pol.act_half.contiguous()}
w = abi.FtlMlpWeights(*[keep[k].data_ptr() for k in ("w1", "b1", "w2", "b2", "w3", "b3", "ns", "mid", "half")], D, A)
act, val = torch.zeros(n, A, device="cuda"), torch.zeros(n, device="cuda")
ptrs = [obs[k].data_ptr() for k in range(4)]
f = L.ftl_policy_mlp; wr = C.byref(w); npx, ap, vp = noise.data_ptr(), act.data_ptr(), val.data_ptr()
for k in range(20): f(wr, ptrs[k % 4], D, npx, n, ap, vp, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(200): f(wr, ptrs[k % 4], D, npx, n, ap, vp, None)
e1.record(); torch.cuda.synchronize()
print("%.2f us per launch   checksum %.6f" % (e0.elapsed_time(e1) * 5.0, float(act.double().sum() + val.double().sum())))
'''
runs = [("default tcgen05", None, None), ("default mma.sync", None, "mma")] + [(os.path.basename(p), os.path.abspath(p), None) for p in sorted(glob.glob("tools/libftl_pol_*.so"))]
for name, lib, impl in runs:
    env = dict(os.environ)
    if lib: env["POL_LIB"] = lib
    if impl: env["FTL_POLICY_IMPL"] = impl
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("%-28s %s" % (name, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]), flush=True)
