"""Two half-batches as a software pipeline with ORDERED launches: the kinematics kernel of one half is enqueued before the
ray kernels of the other half, so that the latency-bound k_kin and the issue-bound k_rays share the SMs for the whole
step (tools/pipeline_ab.py left the order to whole-step calls: every k_kin then queues behind a ray kernel and starves).

    per tick:  SA: k_kin(A_t) | SB: k_rays(B_t-1), k_finish | SB: k_kin(B_t) | SA: k_rays(A_t), k_finish

Diagnostic; prints ms per 65 536 env-steps and a checksum of both halves' states.
"""
import hashlib, os, sys
import torch
sys.path.insert(0, ".")
import bench
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv

OPT_STEP_PHASE, OPT_NO_OVERLAP = 2, 3


def setopt(e, opt, v):
    rc = e._L.ftl_set_option(e._h, opt, v)
    assert rc == 0, rc


def run(mode, n_total=65536, steps=200, settle=150, no_overlap=False, prio=False):
    gc = bench.workload_config(True)
    pool, _ = bench.workload_pool(gc)
    parts = 1 if mode == "whole" else 2
    n = n_total // parts
    envs = [FtlBatchEnv(n, game_config=gc, scenario_pool=pool, env_id_base=k * n, lib_path=os.environ.get("AB_LIB")) for k in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    hi_streams = [torch.cuda.Stream(priority=-1) for _ in range(parts)]   # prio mode: the kinematics kernels run here
    ev_kin = [torch.cuda.Event() for _ in range(parts)]
    ev_rays = [torch.cuda.Event() for _ in range(parts)]
    g = torch.Generator(device="cuda").manual_seed(1234)
    lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
    acts = (lo + (hi - lo) * torch.rand((16, n_total, 2), generator=g, device="cuda")).contiguous()
    views = [[acts[j, k * n:(k + 1) * n] for j in range(16)] for k in range(parts)]
    torch.cuda.synchronize()
    for e, st in zip(envs, streams):
        with torch.cuda.stream(st):
            e.reset()
        if no_overlap:
            setopt(e, OPT_NO_OVERLAP, 1)
    torch.cuda.synchronize()
    if prio:
        for k in range(parts):
            with torch.cuda.stream(streams[k]):
                ev_rays[k].record()
    state = {"t": 0, "primed": False}

    def phase(k, ph):
        e, st = envs[k], streams[k]
        setopt(e, OPT_STEP_PHASE, ph)
        if prio:   # k_kin on a high-priority stream, the ray kernels on a normal one, events between them
            if ph == 1:
                hi_streams[k].wait_event(ev_rays[k])
                with torch.cuda.stream(hi_streams[k]):
                    e.step_raw(views[k][state["t"] % 16])
                    ev_kin[k].record()
            else:
                st.wait_event(ev_kin[k])
                with torch.cuda.stream(st):
                    e.step_raw(views[k][state["t"] % 16])
                    ev_rays[k].record()
            return
        with torch.cuda.stream(st):
            e.step_raw(views[k][state["t"] % 16])

    def go(count):
        for _ in range(count):
            if mode == "whole":
                with torch.cuda.stream(streams[0]):
                    envs[0].step_raw(views[0][state["t"] % 16])
            elif mode == "alternate":     # whole steps, alternating handles (what pipeline_ab.py does)
                for k in range(2):
                    with torch.cuda.stream(streams[k]):
                        envs[k].step_raw(views[k][state["t"] % 16])
            else:                          # ordered halves
                phase(0, 1)                          # k_kin(A_t)
                if state["primed"]: phase(1, 2)      # rays(B_t-1)
                phase(1, 1)                          # k_kin(B_t)
                phase(0, 2)                          # rays(A_t)
                state["primed"] = True
            state["t"] += 1

    def drain():
        if mode == "ordered" and state["primed"]:
            phase(1, 2)
            state["primed"] = False

    go(settle)
    drain()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(parts)]
    e0.record(streams[0])
    for st in streams[1:] + (hi_streams if prio else []):
        st.wait_event(e0)
    go(steps)
    drain()
    for k, (ev, st) in enumerate(zip(ends, streams)):
        if prio: st.wait_event(ev_kin[k])
        ev.record(st)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in ends) / steps
    hh = hashlib.sha256()
    for e in envs:
        setopt(e, OPT_STEP_PHASE, 0)
        s = e.get_state(0, 2048)
        hh.update(bytes(memoryview(s.env)))
        e.close()
    return ms, hh.hexdigest()[:12]


if __name__ == "__main__":
    for name, kw in (("whole 1 x 65536", dict(mode="whole")),
                     ("alternate 2 x 32768 (whole steps)", dict(mode="alternate")),
                     ("ordered halves", dict(mode="ordered")),
                     ("ordered halves, plain stream order in-step", dict(mode="ordered", no_overlap=True)),
                     ("ordered halves, k_kin on high-priority streams", dict(mode="ordered", no_overlap=True, prio=True))):
        ms, h = run(**kw)
        print("%-46s %.4f ms per 65536 env-steps -> %.1f M env-steps/s   %s" % (name, ms, 65536 / ms / 1e3, h), flush=True)
