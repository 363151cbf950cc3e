"""The reference's gym surface on top of the C-ABI (CPU: the host build of the device functions).

Checks the drop-in contract of SURVEY.md section 8(b): ids, constructor keywords and their exceptions,
spaces, seed/reset/step with the 4-tuple API, info strings, the attributes wrappers read, and the two
wrappers of the shipped configurations.
"""
import json

import numpy as np
import pytest

import parity
from continiousenvironment_follower_leader_b200 import gym_surface as gs, scenario_gen, wrappers
from continiousenvironment_follower_leader_b200.config import cfg3_sensors, GameConfig
from hostsim_py import lib


def _make(**kw):
    return gs.make("Test-Cont-Env-Auto-v0", lib=lib(), **kw)


def test_registered_ids_match_the_reference():
    assert set(gs.REGISTERED) == {
        "Test-Cont-Env-Auto-v0", "Test-Cont-Env-Manual-v0", "Test-Cont-Env-Manual-gazebo-v0",
        "Test-Cont-Env-Manual-hardcore-v0", "Test-Cont-Env-Manual-gazebo-hardcore-v0",
        "Test-Cont-Env-Auto-Follow-no-obstacles-v0", "Test-Cont-Env-Auto-Follow-with-obstacles-v0", "Test-Game-Neat-v0"}
    with pytest.raises(AttributeError):          # registered against classes that do not exist upstream either
        gs.make("Test-Cont-Env-Manual-hardcore-v0")
    with pytest.raises(NotImplementedError):     # keyboard teleop presets
        gs.make("Test-Cont-Env-Manual-v0")


def test_spaces_match_the_reference_defaults():
    env = _make()
    np.testing.assert_allclose(env.action_space.low, [0.0, -0.57296], rtol=1e-6)
    np.testing.assert_allclose(env.action_space.high, [0.25, 0.57296], rtol=1e-6)
    assert env.action_space.low.dtype == np.float32
    np.testing.assert_allclose(env.observation_space.high,
                               [1500, 1000, 0.25, 360, 0.57296, 1500, 1000, 0.25, 360, 0.57296], rtol=1e-6)
    assert env.observation_space.low[4] == pytest.approx(-0.57296)
    neg = _make(negative_speed=True, follower_max_speed=0.6)
    np.testing.assert_allclose(neg.action_space.low, [-0.3, -0.57296], rtol=1e-6)   # Env_demo.ipynb cell 7
    assert gs.TestGameNEAT.__mro__[1] is gs.Game
    assert GameConfig(discrete_action_space=True).c.action_mode == 2


def test_constructor_errors_match_the_reference():
    with pytest.raises(ValueError):
        _make(path_finding_algorythm="rrt")
    with pytest.raises(ValueError):
        _make(add_bear=True, bear_number=0)
    with pytest.raises(NotImplementedError):
        _make(multiple_end_points=True, path_finding_algorythm="astar")
    with pytest.raises(ValueError):
        _make(follower_sensors={"Mystery": {"foo": 1}})
    with pytest.raises(AssertionError):          # max_prev_obs must be positive, SEN:876
        _make(follower_sensors=dict(cfg3_sensors(), bad={"sensor_class": "LeaderCorridor_Prev_lasers_v2"}))


@pytest.mark.parametrize("trace", ["cfg3_seed5_follow", "cfg3_seed11_random", "cfg1_auto_seed0_random"])
def test_seeded_reset_draws_the_reference_layout(trace):
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/" + trace + ".npz")
    env = _make(**meta["kwargs"])
    env.seed(meta["seed"])
    env.reset()
    pool = env._env._pool
    ns = int(pool.n_static[0])
    assert np.array_equal(pool.static_rects[0, :ns], d["scen_static_rects"])    # walls + 35 rocks, same RNG draws
    assert np.array_equal(pool.leader_pos[0], d["scen_leader_pos"])
    assert tuple(env.trajectory[0]) == tuple(d["scen_route"][0])
    if bool(d["scen_found_target_point"]):
        assert env.found_target_point
        # same grid, same metric: the planned route has the reference's length (waypoints differ by tie-breaks)
        assert abs(len(env.trajectory) - len(d["scen_route"])) <= 2
    else:
        assert not env.found_target_point        # seed 0 is an unreachable-route seed upstream too


def test_step_returns_the_old_gym_four_tuple_with_reference_types():
    env = _make(follower_sensors=cfg3_sensors(), bear_number=1)
    env.seed(5)
    obs = env.reset()
    assert set(obs) == {"numerical_features", "leader_target_point", "LeaderPositionsTracker_v2",
                        "LeaderCorridor_Prev_lasers_v2", "LaserPrevSensor"}
    assert obs["numerical_features"].shape == (10,) and obs["numerical_features"].dtype == np.float32
    assert obs["LeaderCorridor_Prev_lasers_v2"].shape == (5, 12) and obs["LaserPrevSensor"].shape == (5, 36)
    hist, corridor = obs["LeaderPositionsTracker_v2"]
    assert len(hist) == len(corridor) > 5
    out = env.step(np.array([0.25, 0.0], np.float32))
    assert len(out) == 4
    obs, reward, done, info = out
    assert isinstance(reward, float) and isinstance(done, bool)
    assert info == {"mission_status": "in_progress", "agent_status": "moving", "leader_status": "moving"}
    assert env.step_count == 10 and env.overall_reward == 10 * reward   # ten frames of +1; the step returns the last frame
    assert env.follower.sensors["LaserPrevSensor"].laser_length == 200
    assert env.max_distance == 200 and env.PIXELS_TO_METER == 50
    assert env.follower.position.dtype == np.float32 and len(env.leader_factual_trajectory) > 10


def test_game_replays_a_reference_trace_when_given_its_scenario():
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/cfg3_seed5_follow.npz")
    env = _make(**meta["kwargs"])
    sc = scenario_gen.Scenario()
    sc.static_rects = [tuple(r) for r in d["scen_static_rects"]]
    sc.route = [tuple(p) for p in d["scen_route"]]
    sc.leader_pos, sc.leader_dir = d["scen_leader_pos"], float(d["scen_leader_dir"])
    sc.follower_pos, sc.follower_dir = d["scen_follower_pos"], float(d["scen_follower_dir"])
    sc.found_target_point = True
    obs = env.reset(scenario=sc)
    assert np.array_equal(obs["numerical_features"], d["t_nf"][0])
    for t, a in enumerate(d["actions"][:120]):
        obs, reward, done, info = env.step(a)
        assert np.array_equal(obs["numerical_features"], d["t_nf"][t + 1])
        assert reward == d["t_floats"][t + 1][0] and done == bool(d["t_ints"][t + 1][4])
        got = np.concatenate([obs[n].reshape(-1) for n in meta["ray_names"]])
        assert np.allclose(got, d["t_rays"][t + 1], rtol=parity.RTOL)


def test_sensor_prev_wrapper_and_skip_bad_seeds():
    base = _make(follower_sensors=cfg3_sensors(), bear_number=1)
    env = wrappers.SkipBadSeeds(wrappers.ContinuousObserveModifier_sensorPrev(
        base, action_values_range=[-1, 1], max_prev_obs=5))
    assert env.observation_space.shape == (5, 48)
    assert np.all(env.action_space.low == -1) and np.all(env.action_space.high == 1)
    env.seed(0)                                   # seed 0: unreachable target upstream -> SkipBadSeeds resets again
    obs = env.reset()
    assert base.found_target_point and base.simulation_number >= 2
    assert obs.shape == (5, 48) and obs.min() >= 0 and obs.max() <= 1
    raw, _, _, _ = base.step([0.2, 0.1])
    want = np.concatenate([np.clip(raw["LeaderCorridor_Prev_lasers_v2"] / 150, 0, 1),
                           np.clip(raw["LaserPrevSensor"] / 200, 0, 1)], axis=1)
    assert np.array_equal(env.env.observation(raw), want)
    with pytest.raises(KeyError):                 # every sensor config must carry "sensor_class", WRP:181
        wrappers.ContinuousObserveModifier_sensorPrev(_make(follower_sensors={
            "LeaderPositionsTracker_v2": {k: v for k, v in cfg3_sensors()["LeaderPositionsTracker_v2"].items()
                                          if k != "sensor_class"}}), max_prev_obs=5)


def test_discrete_and_constant_speed_action_spaces():
    env = _make(discrete_action_space=True, add_bear=False)
    env.seed(3)
    env.reset()
    obs, r, d, info = env.step(4)
    assert obs["numerical_features"][9] > 0       # turning at +max rotation speed
    env2 = _make(constant_follower_speed=True, add_bear=False)
    assert env2.action_space.shape == (1,)
    env2.seed(3)
    env2.reset()
    obs, r, d, info = env2.step(np.array([0.0], np.float32))
    assert obs["numerical_features"][7] == pytest.approx(0.025)   # ten frames of 0.0025 px/frame^2


def test_compas_sensor_through_the_gym_surface_and_the_sensor_prev_wrapper():
    """LeaderCorridor_lasers_compas (SEN:1138-1240): (H, 5 * R) blocks in the obs dict, 5 * R features per row in
    ContinuousObserveModifier_sensorPrev (WRP:187-188, 214-219), and the reference's flag check."""
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/compas_seed25.npz")
    env = _make(**meta["kwargs"])
    sc = scenario_gen.Scenario()
    sc.static_rects = [tuple(r) for r in d["scen_static_rects"]]
    sc.route = [tuple(p) for p in d["scen_route"]]
    sc.leader_pos, sc.leader_dir = d["scen_leader_pos"], float(d["scen_leader_dir"])
    sc.follower_pos, sc.follower_dir = d["scen_follower_pos"], float(d["scen_follower_dir"])
    sc.found_target_point = True
    wrapped = wrappers.ContinuousObserveModifier_sensorPrev(env, max_prev_obs=3)
    assert wrapped.observation_space.shape == (3, 5 * 12 + 36)   # WRP:181-188: 5 * R for the compas sensor + the history sensor
    obs = env.reset(scenario=sc)
    for t, a in enumerate(d["actions"][:60]):
        obs, reward, done, info = env.step(a)
        block = obs["LeaderCorridor_lasers_compas"]
        assert block.shape == (3, 60)
        got = np.concatenate([obs[n].reshape(-1) for n in meta["ray_names"]])
        assert np.allclose(got, d["t_rays"][t + 1], rtol=parity.RTOL)
        # every ray reports in exactly one of its five columns
        cols = block.reshape(3, 5, 12)
        assert np.all((cols > 0).sum(axis=1) == 1)
    with pytest.raises(ValueError):
        _make(follower_sensors=dict(cfg3_sensors(), c={"sensor_class": "LeaderCorridor_lasers_compas", "max_prev_obs": 2,
                                                      "react_to_green_zone": False}))


def test_laser_sensor_through_the_gym_surface():
    """LaserSensor (SEN:18-136): 1 + 2 * ceil(180 / 10) = 37 beams of 20 samples; the dict entry is the (37, 2) array of
    vectors from the follower, equal to the reference trace."""
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/laser_sensor_seed27.npz")
    env = _make(**meta["kwargs"])
    sc = scenario_gen.Scenario()
    sc.static_rects = [tuple(r) for r in d["scen_static_rects"]]
    sc.route = [tuple(p) for p in d["scen_route"]]
    sc.leader_pos, sc.leader_dir = d["scen_leader_pos"], float(d["scen_leader_dir"])
    sc.follower_pos, sc.follower_dir = d["scen_follower_pos"], float(d["scen_follower_dir"])
    sc.found_target_point = True
    obs = env.reset(scenario=sc)
    assert obs["LaserSensor"].shape == (37, 2) and np.array_equal(obs["LaserSensor"], d["t_laser"][0])
    for t, a in enumerate(d["actions"][:80]):
        obs, reward, done, info = env.step(a)
        assert np.array_equal(obs["LaserSensor"], d["t_laser"][t + 1]), t
    with pytest.raises(NotImplementedError):
        _make(follower_sensors={"LaserSensor": {"return_all_points": True}})


def test_render_modes():
    """Game.render: only "rgb_array" exists (ftl_render_host, GPU test in test_gpu_parity_gaps.py); "human" needs a pygame
    window and raises, and so does rendering before the first reset."""
    env = _make(follower_sensors=cfg3_sensors(), bear_number=1)
    assert env.metadata["render.modes"] == ["rgb_array"]
    with pytest.raises(NotImplementedError):
        env.render(mode="human")
    with pytest.raises(RuntimeError):
        env.render(mode="rgb_array")


def test_multiple_end_points_through_the_gym_surface():
    """Game(multiple_end_points=True): the three finish points the reference draws for the seed (ENV:471-482; fixture from
    the unmodified reference, oracle/gen_multi_end_golden.py), a route through all of them, and the leader follows it."""
    with open(parity.GOLDEN_DIR + "/multi_end_points.json") as f:
        gold = json.load(f)
    case = gold["cases"][0]
    env = _make(**gold["kwargs"])
    assert env.finish_point2 == (1490, 990)                       # ENV:246
    env.seed(case["seed"])
    env.reset()
    assert [list(env.finish_point), list(env.finish_point2), list(env.finish_point3)] == case["finish_points"]
    assert len(env.trajectory) == case["n_route"] and env.found_target_point
    with pytest.raises(NotImplementedError):                      # ENV:239-243
        _make(multiple_end_points=True, path_finding_algorythm="astar")
    for _ in range(5):
        obs, rew, done, info = env.step(np.array([0.3, 0.0], np.float32))
    assert np.isfinite(obs["numerical_features"]).all() and not done
