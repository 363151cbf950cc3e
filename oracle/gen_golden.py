"""Export golden traces from the UNMODIFIED reference -- TEST INFRASTRUCTURE, build container only.

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Each trace holds (a) the scenario ``Game.reset()`` built (as FtlScenarioPool arrays), (b) the
constructor kwargs, (c) the float32 action fed at every step, and (d) after reset and after every
step: the observation, reward, done, info codes and the simulator's internal state.  The reference
is run through oracle/ref_harness.py (gym/pygame shims, python-float actions, virtual clock).
tests/test_oracle_golden.py replays the actions through oracle/ftl_oracle.c and demands equality;
the GPU parity tests replay them through libftl.so.
"""
import json
import math
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.dirname(_HERE))

import ref_harness as rh  # noqa: E402
from continiousenvironment_follower_leader_b200 import abi  # noqa: E402
from continiousenvironment_follower_leader_b200.config import (  # noqa: E402
    GameConfig, cfg3_sensors, TEST_GAME_MANUAL_GAZEBO_KWARGS)

GOLDEN_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden")

MISSION = {k: i for i, k in enumerate(abi.MISSION_STATUS)}
AGENT = {k: i for i, k in enumerate(abi.AGENT_STATUS)}
LEADER = {k: i for i, k in enumerate(abi.LEADER_STATUS)}


# ---- policies (they only produce the recorded actions; replays never need them) -------------------
def policy_random(rng, env, lo, hi):
    return rng.uniform(lo, hi).astype(np.float32)


def policy_follow(rng, env, lo, hi, lag=70.0, noise=0.05):
    """Pure-pursuit of a trail point ~lag px behind the leader: keeps the follower in the green zone."""
    trail = env.leader_factual_trajectory
    tgt = np.asarray(trail[-1], dtype=np.float64)
    acc = 0.0
    for i in range(len(trail) - 1, 0, -1):
        acc += math.hypot(trail[i][0] - trail[i - 1][0], trail[i][1] - trail[i - 1][1])
        tgt = np.asarray(trail[i - 1], dtype=np.float64)
        if acc >= lag:
            break
    f = env.follower
    d = tgt - f.position
    want = math.degrees(math.atan2(d[1], d[0])) % 360
    err = (want - f.direction + 540) % 360 - 180
    w = float(np.clip(err * 0.2, lo[1], hi[1]))
    gap = float(np.linalg.norm(np.asarray(env.leader.position, np.float64) - f.position))
    v = hi[0] if gap > 75 else hi[0] * 0.5 if gap > 60 else 0.0
    a = np.array([v, w], dtype=np.float64) + rng.normal(0, noise, 2) * (hi - lo)
    return np.clip(a, lo, hi).astype(np.float32)


def policy_straight(rng, env, lo, hi):
    return np.array([hi[0], 0.0], np.float32)


POLICIES = {"random": policy_random, "follow": policy_follow, "straight": policy_straight}


# ---- state capture ----------------------------------------------------------------------------------
def _robot(r):
    rect = r.rectangle
    return ([float(r.position[0]), float(r.position[1]), float(r.direction), float(r.speed),
             float(r.rotation_speed), float(r.desirable_speed), float(r.desirable_rotation_speed)],
            [int(r.rotation_direction), int(r.desirable_rotation_direction), rect.x, rect.y, rect.w, rect.h])


def capture(env, obs, reward, done, info, ray_names, n_bears, gc=None):
    rec = {}
    robots = [env.follower, env.leader] + list(env.game_dynamic_list)[:n_bears]
    rf, ri = zip(*[_robot(r) for r in robots])
    rec["robot_f"] = np.array(rf, np.float64)          # [2+B, 7]
    rec["robot_i"] = np.array(ri, np.int32)            # [2+B, 6]
    if n_bears:
        rec["bear_target"] = np.array([[float(p[0]), float(p[1])] for p in env.cur_points_for_bear[:n_bears]], np.float64)
        rec["bear_index"] = np.array(env.dynamics_index[:n_bears], np.int32)
    ft = env.finish_position_framestimer
    rec["ints"] = np.array([env.step_count, env.cur_target_id, int(env.leader_finished), -1 if ft is None else ft,
                            int(env.done), int(env.crash), int(env.is_in_box), int(env.is_on_trace),
                            int(env.follower_too_close), len(env.leader_factual_trajectory),
                            MISSION[info["mission_status"]], AGENT[info["agent_status"]],
                            LEADER[info["leader_status"]], len(env.green_zone_trajectory_points)], np.int32)
    rec["floats"] = np.array([float(reward), float(env.overall_reward), float(env.accumulated_penalty),
                              float(env.cur_speed_multiplier)], np.float64)
    rec["trail_last"] = np.array(env.leader_factual_trajectory[-1], np.float32)
    rec["nf"] = np.asarray(obs["numerical_features"], np.float32)
    rec["target"] = np.array(obs["leader_target_point"], np.int32)
    tr = env.follower.sensors.get("LeaderPositionsTracker_v2")
    if tr is not None:
        hist = list(tr.leader_positions_hist)
        corr = list(tr.corridor)
        rec["tracker_i"] = np.array([tr.saving_counter, len(hist), len(corr),
                                     sum(1 for h in hist if h.dtype == np.float64)], np.int32)
        rec["hist_last"] = np.array(hist[-1], np.float64)
        rec["hist_first"] = np.array(hist[0], np.float64)
        rec["corr_last"] = np.array([corr[-1][0], corr[-1][1]], np.float64).reshape(4)
        rec["corr_first"] = np.array([corr[0][0], corr[0][1]], np.float64).reshape(4)
    if ray_names:
        rec["rays"] = np.concatenate([np.asarray(obs[n], np.float32).reshape(-1) for n in ray_names])
    if gc is not None and gc.follower_info_name is not None:
        rec["follower_info"] = np.asarray(obs[gc.follower_info_name], np.float32)
    if gc is not None and gc.track_vector_name is not None:
        rec["track_vectors"] = np.asarray(obs[gc.track_vector_name], np.float32)
    if gc is not None and getattr(gc, "radar_name", None) is not None:
        rec["radar"] = np.asarray(obs[gc.radar_name], np.float32)
    if gc is not None and getattr(gc, "laser_name", None) is not None:
        rec["laser"] = np.asarray(obs[gc.laser_name], np.float32)
    return rec


def extract_scenario(env):
    return dict(
        static_rects=rh.static_rects(env),
        route=np.array(env.trajectory, np.int32).reshape(-1, 2),
        leader_pos=np.array(env.leader.position, np.float32),
        leader_dir=np.float64(env.leader.direction),
        follower_pos=np.array(env.follower.position, np.float32),
        follower_dir=np.float64(env.follower.direction),
        found_target_point=np.uint8(bool(env.found_target_point)),
    )


def full_state_tail(env):
    """Variable-length internals, stored once at the end of a trace for a deep check."""
    out = {"trail": np.array([[p[0], p[1]] for p in env.leader_factual_trajectory], np.float32)}
    tr = env.follower.sensors.get("LeaderPositionsTracker_v2")
    if tr is not None:
        out["hist"] = np.array([np.asarray(h, np.float64) for h in tr.leader_positions_hist], np.float64)
        out["corridor"] = np.array([[c[0][0], c[0][1], c[1][0], c[1][1]] for c in tr.corridor], np.float64)
    return out


def run_trace(name, env_id, kwargs, seed, policy, max_env_steps, action_seed=0, until_done=False, switch=None):
    """switch: optional (step, policy) to change policy mid-episode."""
    t0 = time.time()
    env = rh.make_env(env_id, **kwargs)
    obs = rh.reset(env, seed=seed)
    gc = GameConfig(**_config_kwargs(env_id, kwargs))
    ray_names = gc.ray_sensor_names
    n_bears = gc.c.n_bears
    lo, hi = gc.action_bounds()
    lo, hi = lo.astype(np.float64), hi.astype(np.float64)
    rng = np.random.RandomState(action_seed)
    scen = extract_scenario(env)
    info0 = {"mission_status": "in_progress", "agent_status": "moving", "leader_status": "moving"}
    recs = [capture(env, obs, 0.0, False, info0, ray_names, n_bears, gc)]
    actions = []
    pol = POLICIES[policy]
    # traces whose steps consume the reference's global RNGs also store what was drawn (FtlStepInputs of a replay)
    list_regime = any(isinstance(v, (list, tuple)) for v in (kwargs.get("leader_speed_regime") or {}).values())
    recorder = rh.StepInputRecorder(env, gc.c.frames_per_step) if (list_regime or gc.random_frames_per_step) else None
    for t in range(max_env_steps):
        if switch is not None and t == switch[0]:
            pol = POLICIES[switch[1]]
        a = pol(rng, env, lo, hi)
        if recorder is not None:
            recorder.begin_step()
        obs, reward, done, info = rh.step(env, a)
        if recorder is not None:
            recorder.end_step()
        actions.append(a)
        recs.append(capture(env, obs, reward, done, info, ray_names, n_bears, gc))
        if until_done and done:
            break
    out = {"actions": np.array(actions, np.float32)}
    if recorder is not None:
        out["step_frames"] = np.array(recorder.frames, np.int32)
        out["step_draws"] = np.array(recorder.draws, np.float64)
    for k in recs[0]:
        out["t_" + k] = np.stack([r[k] for r in recs])
    for k, v in scen.items():
        out["scen_" + k] = v
    for k, v in full_state_tail(env).items():
        out["final_" + k] = v
    out["meta"] = np.array(json.dumps(dict(name=name, env_id=env_id, kwargs=kwargs, seed=seed, policy=policy,
                                            ray_names=ray_names, n_env_steps=len(actions),
                                            numpy=np.__version__)))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    last = recs[-1]["ints"]
    print("%-28s steps=%4d frames=%5d done=%d crash=%d mission=%s overall=%.1f  (%.1fs, %d KB)" % (
        name, len(actions), last[0], last[4], last[5], abi.MISSION_STATUS[last[10]], recs[-1]["floats"][1],
        time.time() - t0, os.path.getsize(path) // 1024))
    return path


def _config_kwargs(env_id, kwargs):
    """TestGameAuto forwards its kwargs unchanged (ENV:1963-1965).  The "gazebo" preset
    (TestGameManual_gazebo, ENV:2013-2107) hard-codes manual_control=True, so its physics is reached
    by passing the same kwargs to TestGameAuto."""
    assert env_id == "Test-Cont-Env-Auto-v0"
    return dict(kwargs)


TRACES = [
    # BASELINE.json configs[0]: default env, seed 0 (an unreachable-route seed: 15000-point looped route)
    dict(name="cfg1_auto_seed0_random", env_id="Test-Cont-Env-Auto-v0", kwargs={}, seed=0, policy="random",
         max_env_steps=520, until_done=True),
    # default env, a good seed, follower keeps to the trail until the leader finishes (success path)
    dict(name="auto_nobear_seed5_follow", env_id="Test-Cont-Env-Auto-v0", kwargs=dict(add_bear=False), seed=5,
         policy="follow", max_env_steps=520, until_done=True),
    # configs[2] shape: 35 rocks + 1 bear, tracker + 12-ray corridor sensor + 36-ray obstacle sensor
    dict(name="cfg3_seed5_follow", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(follower_sensors=cfg3_sensors(), bear_number=1), seed=5, policy="follow", max_env_steps=300),
    dict(name="cfg3_seed11_random", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(follower_sensors=cfg3_sensors(), bear_number=1), seed=11, policy="random", max_env_steps=200,
         until_done=True),
    dict(name="cfg3_seed23_follow_then_random", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(follower_sensors=cfg3_sensors(), bear_number=1), seed=23, policy="follow", max_env_steps=260,
         switch=(150, "random")),
    # configs[1] shape: no obstacles, no bears, tracker only (A* route; D* needs the bridge objects, ENV:1501)
    dict(name="cfg2_noobst_seed3_follow", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(add_obstacles=False, add_bear=False, path_finding_algorythm="astar",
                     follower_sensors={"LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"]}),
         seed=3, policy="follow", max_env_steps=250),
    # pad_sectors=True layout, discrete-free variants of the sensor flags, 3 bears, F=5
    dict(name="pad_sectors_seed7", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(frames_per_step=5, bear_number=3, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "LeaderCorridor_Prev_lasers_v2": dict(cfg3_sensors()["LeaderCorridor_Prev_lasers_v2"], pad_sectors=True,
                                                   max_prev_obs=3, react_to_obstacles="dynamic"),
             "static_only": dict(cfg3_sensors()["LaserPrevSensor"], sensor_name="static_only", lasers_count=24,
                                 react_to_obstacles="static", max_prev_obs=2)}),
         seed=7, policy="follow", max_env_steps=200),
    # configs[3] physics ("gazebo" preset, px/m=10, F=5, early stopping, scalar regimes only: the
    # list-valued regime entries draw from python's global MT19937 and cannot be replayed)
    dict(name="gazebo_scalar_regimes_seed2", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, max_steps=4000,
                     leader_speed_regime={0: 0.6, 300: 1, 900: 0.75, 1400: 0, 1600: 1},
                     leader_acceleration_regime={0: 0, 1100: 0.03, 1300: 0}),
         seed=2, policy="follow", max_env_steps=500, until_done=True),
    # early stopping by low reward / too far: follower stands still
    dict(name="gazebo_early_stop_seed4", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, leader_speed_regime={0: 1}, leader_acceleration_regime=None,
                     early_stopping={"max_distance_coef": 1.5, "low_reward": -40}),
         seed=4, policy="random", max_env_steps=400, until_done=True),
    # configs[3] as the preset ships it: list-valued speed regimes (random.uniform per frame, ENV:1155-1156) and
    # random_frames_per_step (np.random.randint per step, ENV:939-940); the draws are stored with the trace
    dict(name="gazebo_list_regimes_random_frames_seed3", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, max_steps=3000, random_frames_per_step=[2, 8],
                     leader_speed_regime={0: [0.2, 1], 100: 1, 180: [0.5, 1], 260: 0.75, 330: [0.0, 0.5], 400: [0.4, 1]},
                     leader_acceleration_regime={0: 0, 200: 0.03, 300: 0}),
         seed=3, policy="follow", max_env_steps=700, until_done=True),
    # LaserSensor (SEN:18-136): the point-sampling lidar, default fan (37 beams) and a narrow distance-only fan, 3 bears
    dict(name="laser_sensor_seed27", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(bear_number=3, follower_sensors={
             "LaserSensor": dict(sensor_class="LaserSensor", available_angle=360, angle_step=10, points_number=20,
                                 sensor_range=5)}),
         seed=27, policy="random", max_env_steps=150),
    dict(name="laser_distances_seed29", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(bear_number=1, frames_per_step=5, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "lidar": dict(sensor_class="LaserSensor", available_angle=150, angle_step=7.5, points_number=30,
                           sensor_range=6, return_only_distances=True)}),
         seed=29, policy="follow", max_env_steps=150),
    # LeaderCorridor_lasers_compas (SEN:1138-1240) next to a history sensor: 5 * R columns per row
    dict(name="compas_seed25", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(bear_number=1, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "LeaderCorridor_lasers_compas": dict(sensor_class="LeaderCorridor_lasers_compas", react_to_safe_corridor=True,
                                                  react_to_green_zone=True, react_to_obstacles=False, lasers_count=12,
                                                  laser_length=150, max_prev_obs=3, use_prev_obs=True, pad_sectors=False),
             "LaserPrevSensor": cfg3_sensors()["LaserPrevSensor"]}),
         seed=25, policy="follow", max_env_steps=220, switch=(150, "random")),
    # the sensors without history on the same ray engine (LeaderCorridor_lasers_v2 = rays at k*360/R,
    # LeaderCorridor_lasers = the fixed 7-ray fan), FollowerInfo and LeaderTrackDetector_vector ("new")
    dict(name="flat_sensors_seed9", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(bear_number=2, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "LeaderCorridor_lasers_v2": dict(sensor_class="LeaderCorridor_lasers_v2", react_to_safe_corridor=True,
                                              react_to_obstacles=True, react_to_green_zone=True, lasers_count=24,
                                              laser_length=180),
             "LeaderCorridor_lasers": dict(sensor_class="LeaderCorridor_lasers", react_to_safe_corridor=True,
                                           react_to_obstacles="dynamic", front_lasers_count=5, back_lasers_count=2,
                                           laser_length=120),
             "FollowerInfo": dict(sensor_class="FollowerInfo"),
             "LeaderTrackDetector_vector": dict(sensor_class="LeaderTrackDetector_vector", position_sequence_length=12,
                                                detectable_positions="new")}),
         seed=9, policy="follow", max_env_steps=220, switch=(160, "random")),
    # the "old" end of the history, a history sensor next to a flat one, 3-ray fan
    dict(name="flat_sensors_old_seed13", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(bear_number=1, frames_per_step=5, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "LeaderCorridor_Prev_lasers_v2": cfg3_sensors()["LeaderCorridor_Prev_lasers_v2"],
             "fan": dict(sensor_class="LeaderCorridor_lasers", react_to_safe_corridor=True, react_to_obstacles=True,
                         laser_length=100),
             "LeaderTrackDetector_vector": dict(sensor_class="LeaderTrackDetector_vector", position_sequence_length=50,
                                                detectable_positions="old")}),
         seed=13, policy="follow", max_env_steps=200),
    # LeaderTrackDetector_radar (SEN:394-461): the oldest 30 points in 12 sectors next to the history ray sensor ...
    dict(name="radar_old_seed17", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(bear_number=1, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "LeaderCorridor_Prev_lasers_v2": cfg3_sensors()["LeaderCorridor_Prev_lasers_v2"],
             "LeaderTrackDetector_radar": dict(sensor_class="LeaderTrackDetector_radar", position_sequence_length=30,
                                               detectable_positions="old", radar_sectors_number=12)}),
         seed=17, policy="follow", max_env_steps=260, switch=(180, "random")),
    # ... and the newest 100 points in 180 sectors (the class defaults apart from "new"), random walk
    dict(name="radar_new_seed19", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(add_bear=False, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "LeaderTrackDetector_radar": dict(sensor_class="LeaderTrackDetector_radar", detectable_positions="new")}),
         seed=19, policy="random", max_env_steps=200),
    # "near": all points, few wide sectors
    dict(name="radar_near_seed21", env_id="Test-Cont-Env-Auto-v0",
         kwargs=dict(add_bear=False, follower_sensors={
             "LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"],
             "radar": dict(sensor_class="LeaderTrackDetector_radar", detectable_positions="near", radar_sectors_number=7)}),
         seed=21, policy="follow", max_env_steps=150),
]


def main(argv):
    only = set(argv[1:])
    for t in TRACES:
        if only and t["name"] not in only:
            continue
        run_trace(**t)


if __name__ == "__main__":
    main(sys.argv)
