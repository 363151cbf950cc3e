set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.txt 2>&1; tail -8 gpurun_out/r2v_tests.txt
timeout 120 python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2v_smoke.txt 2>&1; tail -2 gpurun_out/r2v_smoke.txt
timeout 120 python - > gpurun_out/r2v_render_time.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
gc = bench.workload_config(True); pool, _ = bench.workload_pool(gc)
env = FtlBatchEnv(4096, game_config=gc, scenario_pool=pool); env.reset()
g = torch.Generator(device='cuda').manual_seed(1)
lo, hi = [torch.tensor(x, device='cuda') for x in gc.action_bounds()]
acts = lo + (hi - lo) * torch.rand((8, 4096, 2), generator=g, device='cuda')
for k in range(200): env.step_raw(acts[k % 8])
for n, scale in ((1, 1), (64, 4), (256, 8)):
    env.render(0, n, scale); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): img = env.render(0, n, scale)
    e1.record(); torch.cuda.synchronize()
    print('render %d envs at scale %d: %s, %.3f ms per call' % (n, scale, tuple(img.shape), e0.elapsed_time(e1) / 10))
import numpy as np
np.save('gpurun_out/r2v_frame.npy', env.render(0, 1, 2).cpu().numpy())
PY
cat gpurun_out/r2v_render_time.txt
