"""Device functions (host build) against the oracle on seeded synthetic batches, on the CPU: the same
checks tests/test_gpu_parity.py applies to the CUDA build, at sizes that finish in seconds."""
import numpy as np
import pytest

import parity
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors, TEST_GAME_MANUAL_GAZEBO_KWARGS
from continiousenvironment_follower_leader_b200.scenario import synthetic_pool
from hostsim_py import make_env
from oracle_py import OracleEnv
from test_gpu_parity import _compare_states, _ray_outliers

FLAT_SENSOR_KWARGS = parity.load_trace(parity.GOLDEN_DIR + "/flat_sensors_seed9.npz")[1]["kwargs"]

CASES = [
    ("cfg2", dict(add_obstacles=False, add_bear=False,
                  follower_sensors={"LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"]}), 96, 60),
    ("cfg3", dict(bear_number=1, follower_sensors=cfg3_sensors()), 96, 80),
    ("cfg3_3bears_discrete", dict(bear_number=3, discrete_action_space=True, follower_sensors=cfg3_sensors(24, 20, 3)), 48, 50),
    ("cfg3_autoreset", dict(bear_number=1, follower_sensors=cfg3_sensors(), max_steps=200, auto_reset=True), 64, 70),
    ("gazebo_ranges", dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, max_steps=600, auto_reset=True), 48, 150),
    # BASELINE.json configs[4]: ray-count sweep up to 360 lasers (the reference only accepts 12/20/24/36, SEN:761)
    ("rays_360", dict(bear_number=2, frames_per_step=3, follower_sensors=cfg3_sensors(72, 360, 4)), 24, 40),
    ("rays_120_f1", dict(bear_number=1, frames_per_step=2, follower_sensors=cfg3_sensors(20, 120, 8)), 24, 60),
    # more frames per step than one chunk of the two-pass frame loop (kFrameChunk = 10)
    ("cfg3_f13", dict(bear_number=2, frames_per_step=13, follower_sensors=cfg3_sensors(), max_steps=400, auto_reset=True), 48, 50),
    # sensors without history, FollowerInfo, LeaderTrackDetector_vector (kwargs of the golden trace)
    ("flat_sensors", dict(FLAT_SENSOR_KWARGS, auto_reset=True, max_steps=300), 48, 80),
    # LeaderTrackDetector_radar in its three modes (SEN:394-461)
    ("radar_old", dict(parity.load_trace(parity.GOLDEN_DIR + "/radar_old_seed17.npz")[1]["kwargs"], auto_reset=True,
                       max_steps=300), 48, 70),
    ("radar_new", dict(parity.load_trace(parity.GOLDEN_DIR + "/radar_new_seed19.npz")[1]["kwargs"]), 32, 60),
    ("radar_near", dict(parity.load_trace(parity.GOLDEN_DIR + "/radar_near_seed21.npz")[1]["kwargs"]), 32, 60),
    # the remaining action decode and reward options of Game.step (ENV:918-925, 1136) and the shortest step
    ("const_speed", dict(bear_number=1, constant_follower_speed=True, follower_sensors=cfg3_sensors()), 48, 60),
    ("aggregate_reward", dict(bear_number=1, aggregate_reward=True, follower_sensors=cfg3_sensors(), max_steps=300,
                              auto_reset=True), 48, 60),
    ("f1", dict(bear_number=1, frames_per_step=1, follower_sensors=cfg3_sensors()), 48, 120),
    # LeaderCorridor_lasers_compas (SEN:1138-1240), raw and as part of the fused sensorPrev matrix
    ("compas", dict(parity.load_trace(parity.GOLDEN_DIR + "/compas_seed25.npz")[1]["kwargs"], auto_reset=True,
                    max_steps=300), 48, 80),
    # LaserSensor (SEN:18-136): points and distances
    ("laser_points", dict(parity.load_trace(parity.GOLDEN_DIR + "/laser_sensor_seed27.npz")[1]["kwargs"], auto_reset=True,
                          max_steps=300), 48, 60),
    ("laser_distances", dict(parity.load_trace(parity.GOLDEN_DIR + "/laser_distances_seed29.npz")[1]["kwargs"],
                             auto_reset=True, max_steps=300), 48, 60),
]


@pytest.mark.parametrize("name,kwargs,n,steps", CASES, ids=[c[0] for c in CASES])
def test_hostsim_matches_oracle(name, kwargs, n, steps):
    gc = GameConfig(**kwargs)
    pool = synthetic_pool(gc, 24, seed=1)
    sim, orc = make_env(gc, n), OracleEnv(gc, n, n_threads=4)
    sim.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    sim.reset(scenario_ids=ids)
    orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(7)
    bounds = gc.action_bounds()
    bad, total = 0, 0
    for t in range(steps):
        a = parity.sample_actions(gc, rng, n, t)
        os_, oo = sim.step(a), orc.step(a)
        assert np.array_equal(os_.done, oo.done), "done differs at step %d" % t
        assert np.array_equal(os_.status, oo.status)
        assert np.array_equal(os_.reward, oo.reward)
        assert np.array_equal(os_.leader_target, oo.leader_target)
        assert np.array_equal(os_.numerical_features, oo.numerical_features)
        if gc.rays_per_env:
            bad += _ray_outliers(os_.rays, oo.rays)
            total += os_.rays.size
        if os_.follower_info is not None:
            assert np.array_equal(os_.follower_info, oo.follower_info)
        if os_.track_vectors is not None:
            assert np.array_equal(os_.track_vectors, oo.track_vectors)
        if os_.radar is not None:
            assert np.array_equal(os_.radar, oo.radar), "radar differs at step %d" % t
        if os_.laser is not None:
            assert np.array_equal(os_.laser, oo.laser), "LaserSensor differs at step %d" % t
        _compare_states(sim.get_state(), orc.get_state(), gc, n, 0.0)
    assert bad == 0, "%d of %d ray values outside tolerance" % (bad, total)
    assert int(sim.get_state().env["overflow"].max()) == 0


@pytest.mark.parametrize("libname", ["libftl_hostsim.so", "libftl_hostsim_smallcaps.so"])
def test_grazing_rays_resolve_like_the_reference(libname):
    """Follower parked on integer coordinates with an axis-aligned heading: rays run exactly through rectangle
    corners and along edges, where the float32 predicates are inconclusive and the float64 fallback must
    reproduce the reference's strict ccw decisions."""
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(12, 36, 5))
    pool = synthetic_pool(gc, 8, seed=2)
    n = 64
    from hostsim_py import lib
    from continiousenvironment_follower_leader_b200 import capi
    # the second library has 2 exact-pass records per env: most of these envs overflow and are recast exactly
    sim, orc = capi.HostEnv(gc, n, lib=lib(libname)), OracleEnv(gc, n)
    sim.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    sim.reset(scenario_ids=ids)
    orc.reset(scenario_ids=ids)
    st = orc.get_state()
    rng = np.random.RandomState(3)
    for i in range(n):   # park next to a rock: corner-aligned, edge-aligned or 25 px off, heading a multiple of 45 degrees
        rects = pool.static_rects[ids[i], 2:pool.n_static[ids[i]]]
        x, y, w, h = rects[rng.randint(len(rects))]
        dx, dy = [(-60, 0), (-60, 25), (-60, h), (w + 40, -30), (w // 2, -50), (-40, -40)][i % 6]
        st.env["follower"]["pos"][i] = (x + dx, y + dy)
        st.env["follower"]["dir"][i] = 45.0 * (i % 8)
        st.env["follower"]["speed"][i] = 0.0
        st.env["follower"]["rot_speed"][i] = 0.0
    orc.set_state(st)
    sim.set_state(st)
    zero = np.zeros((n, 2), np.float32)
    for t in range(6):
        a, b = sim.step(zero), orc.step(zero)
        assert np.array_equal(a.numerical_features, b.numerical_features)
        assert _ray_outliers(a.rays, b.rays, rtol=1e-5) == 0


@pytest.mark.parametrize("pad", [False, True], ids=["plain", "pad_sectors"])
def test_fused_sensor_prev_output_is_the_wrapper_applied_to_the_oracle(pad):
    """fused_sensor_prev=True: the ray kernel writes ContinuousObserveModifier_sensorPrev's matrix (WRP:203-221);
    checked bit for bit against the wrapper's arithmetic applied to the same build's raw sensor output, and within
    the ray tolerance against the wrapper applied to the oracle's."""
    sensors = cfg3_sensors(12, 24, 4)
    for name in sensors:
        if "pad_sectors" in sensors[name]:
            sensors[name]["pad_sectors"] = pad
    kwargs = dict(bear_number=1, follower_sensors=sensors)
    n, steps = 48, 40
    gc, gc_raw = GameConfig(fused_sensor_prev=True, **kwargs), GameConfig(**kwargs)
    assert gc.rays_per_env == gc_raw.rays_per_env
    pool = synthetic_pool(gc, 16, seed=3)
    sim, sim_raw, orc = make_env(gc, n), make_env(gc_raw, n), OracleEnv(gc_raw, n, n_threads=4)
    for e in (sim, sim_raw, orc):
        e.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    o_s, o_r, o_o = sim.reset(scenario_ids=ids), sim_raw.reset(scenario_ids=ids), orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(11)
    bounds = gc.action_bounds()
    for t in range(steps):
        same_build = parity.sensor_prev_expected(gc_raw, o_r.rays)
        want = parity.sensor_prev_expected(gc_raw, o_o.rays)
        got = o_s.rays.reshape(want.shape)
        assert same_build.dtype == np.float32 and np.array_equal(got, same_build), "fused output differs at step %d" % t
        assert np.all(np.abs(got - want) <= parity.RTOL), "fused output differs from the oracle's at step %d" % t
        assert got.min() >= 0.0 and got.max() <= 1.0
        a = rng.uniform(bounds[0], bounds[1], size=(n, 2)).astype(np.float32)
        a[: n // 2, 0], a[: n // 2, 1] = bounds[1][0], 0.0
        o_s, o_r, o_o = sim.step(a), sim_raw.step(a), orc.step(a)


def test_per_step_inputs_random_frames_and_regime_draws():
    """FtlStepInputs (include/ftl.h) through the host build: per-env frames per step and caller-supplied regime draws."""
    from test_gpu_parity_gaps import run_per_step_inputs_case
    run_per_step_inputs_case(make_env, 64, 150, 0.0)


def test_routes_through_three_finish_points_are_followed_alike():
    """multiple_end_points (ENV:471-482, 1552-1611): scenarios from the native generator -- three chained D* legs, ~150-270
    waypoints, route_cap 512 -- stepped by the device functions and by the oracle: the leader's waypoint logic over a route
    several times the usual length."""
    from continiousenvironment_follower_leader_b200 import scenario_gen
    gc = GameConfig(multiple_end_points=True, bear_number=1, follower_sensors=cfg3_sensors(), max_steps=1500)
    pool = scenario_gen.generate_pool_native(gc, [1, 2, 3, 4, 6, 8])
    assert gc.c.route_cap == 512 and int(pool.n_route.max()) > 128 and bool(pool.found_target_point.all())
    n, steps = 24, 260
    sim, orc = make_env(gc, n), OracleEnv(gc, n, n_threads=4)
    sim.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    sim.reset(scenario_ids=ids)
    orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(11)
    lo, hi = gc.action_bounds()
    bad = total = 0
    for t in range(steps):
        a = rng.uniform(lo, hi, size=(n, 2)).astype(np.float32)
        a[: n // 2] = (0.6 * hi[0], 0.0)      # half of the batch just drives on: its leaders get far along their routes
        os_, oo = sim.step(a), orc.step(a)
        assert np.array_equal(os_.done, oo.done), "done differs at step %d" % t
        assert np.array_equal(os_.status, oo.status)
        assert np.array_equal(os_.leader_target, oo.leader_target)
        assert np.array_equal(os_.numerical_features, oo.numerical_features)
        bad += _ray_outliers(os_.rays, oo.rays)
        total += os_.rays.size
    assert bad == 0, "%d of %d ray values outside tolerance" % (bad, total)
    st = orc.get_state()
    _compare_states(sim.get_state(), st, gc, n, 0.0)
    assert int(st.env["cur_target_id"].max()) > 60      # leaders are well into their routes
