"""Kernel LOGIC against the reference's golden traces, on the CPU.

tests/hostsim compiles the same host+device functions the CUDA kernels call (csrc/ftl_device.cuh,
ftl_step.cuh, ftl_rays.cuh).  This checks, without a GPU, that the restructured algorithm (cached
green-zone bounds, per-step static pre-filter, de-duplicated ray casting, by-reference history) gives
the reference's results: every integer and every float of the state bit for bit, rays within the
1e-4 contract.  The CUDA build of the same code is checked on the GPU by tests/test_gpu_parity.py.
"""
import pytest

import parity
from hostsim_py import make_env

FILES = parity.golden_files()


@pytest.mark.parametrize("path", FILES, ids=[p.split("/")[-1][:-4] for p in FILES])
def test_device_functions_reproduce_reference_trace(path):
    d, meta = parity.load_trace(path)
    gc = parity.config_for(meta, _route_len=len(d["scen_route"]), _n_static=len(d["scen_static_rects"]))
    env = make_env(gc, 1)
    env.upload_scenarios(parity.pool_for(d, gc))
    T, outliers = parity.replay(env, d, gc, float_rtol=0.0, ray_rtol=parity.RTOL, ray_outlier_budget=0)
    assert T == meta["n_env_steps"]
    parity.check_final_arrays(env.get_state(), d, gc)
    assert int(env.get_state().env[0]["overflow"]) == 0


def test_ray_pass_overflow_paths_give_the_same_rays():
    """Tiny shared-memory lists (6 edges, 5 pairs) force every overflow branch of the ray pass."""
    from hostsim_py import lib
    from continiousenvironment_follower_leader_b200 import capi
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/cfg3_seed23_follow_then_random.npz")
    gc = parity.config_for(meta, _route_len=len(d["scen_route"]), _n_static=len(d["scen_static_rects"]))
    env = capi.HostEnv(gc, 1, lib=lib("libftl_hostsim_smallcaps.so"))
    env.upload_scenarios(parity.pool_for(d, gc))
    parity.replay(env, d, gc, float_rtol=0.0, ray_rtol=parity.RTOL, ray_outlier_budget=0, max_steps=120)


@pytest.mark.parametrize("libname", ["libftl_hostsim_lanes16.so", "libftl_hostsim_order.so"])
def test_ray_pass_variants_give_the_same_rays(libname):
    """FTL_RAYS_LANES=16 (two envs per warp on the GPU: the lane phases with 16 lanes, smaller lists, the range ring
    handled by lanes 8-15) and a 24-edge list served in a permuted lane order -- same rays as the reference trace."""
    from hostsim_py import lib
    from continiousenvironment_follower_leader_b200 import capi
    for name in ("cfg3_seed23_follow_then_random", "flat_sensors_seed9"):
        d, meta = parity.load_trace(parity.GOLDEN_DIR + "/" + name + ".npz")
        gc = parity.config_for(meta, _route_len=len(d["scen_route"]), _n_static=len(d["scen_static_rects"]))
        env = capi.HostEnv(gc, 1, lib=lib(libname))
        env.upload_scenarios(parity.pool_for(d, gc))
        parity.replay(env, d, gc, float_rtol=0.0, ray_rtol=parity.RTOL, ray_outlier_budget=0, max_steps=150)
