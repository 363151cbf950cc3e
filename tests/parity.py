"""Shared trace-replay checker: the same comparisons are applied to the CPU oracle
(tests/test_oracle_golden.py) and to the CUDA path (tests/test_gpu_parity.py).

A "stepper" is anything with reset(scenario_ids=...) -> outputs, step(actions) -> outputs and
get_state() -> HostState-like (numpy views laid out as include/ftl.h), for one or more envs.
"""
import glob
import json
import os

import numpy as np

from continiousenvironment_follower_leader_b200 import abi
from continiousenvironment_follower_leader_b200.config import GameConfig
from continiousenvironment_follower_leader_b200.scenario import ScenarioPool

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# float tolerance of the parity contract (BASELINE.json north_star): 1e-4 relative
RTOL = 1e-4


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_trace(path):
    d = np.load(path, allow_pickle=False)
    meta = json.loads(str(d["meta"]))
    return d, meta


def config_for(meta, **overrides):
    kw = dict(meta["kwargs"])
    # JSON turns int dict keys into strings; the reference converts regime keys with int() anyway
    route_len = overrides.pop("_route_len", None)
    n_static = overrides.pop("_n_static", None)
    if route_len is not None:
        overrides.setdefault("route_cap", int(route_len))
    if n_static is not None:
        overrides.setdefault("static_cap", max(int(n_static), 1))
    return GameConfig(**kw, **overrides)


def pool_for(d, gc, copies=1):
    pool = ScenarioPool(copies, gc.c.static_cap, gc.c.route_cap)
    for i in range(copies):
        pool.set(i, d["scen_static_rects"], d["scen_route"], d["scen_leader_pos"], float(d["scen_leader_dir"]),
                 d["scen_follower_pos"], float(d["scen_follower_dir"]), bool(d["scen_found_target_point"]))
    return pool


def _robots(env_rec, n_bears):
    return [env_rec["follower"], env_rec["leader"]] + [env_rec["bear"][b] for b in range(n_bears)]


class Mismatch(AssertionError):
    pass


def _close(a, b, rtol, what, t):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b)
    tol = rtol * np.maximum(1.0, np.abs(b))
    if not np.all(err <= tol):
        i = int(np.argmax(err - tol))
        raise Mismatch("step %d: %s differs: got %r want %r (|err| %.3g)" % (
            t, what, a.reshape(-1)[i], b.reshape(-1)[i], err.reshape(-1)[i]))


def _equal(a, b, what, t):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        raise Mismatch("step %d: %s differs: got %r want %r" % (t, what, a.tolist(), b.tolist()))


def check_record(t, d, gc, out, st, env_index=0, float_rtol=0.0, ray_rtol=RTOL, ray_outlier_budget=None):
    """Compare golden record t with the stepper's outputs/state for env `env_index`.

    float_rtol = 0 demands bit-equality of every float (the oracle's bar); the CUDA path is held to
    RTOL for floats and to equality for every integer.  Returns the number of ray values outside
    ray_rtol (0 unless ray_outlier_budget allows some)."""
    c = gc.c
    e = st.env[env_index]
    nb = c.n_bears
    ints = d["t_ints"][t]
    # ---- integers: bit exact -----------------------------------------------------------------
    got_ints = [e["step_count"], e["cur_target_id"], e["leader_finished"], e["finish_timer"], e["done"], e["crash"],
                e["is_in_box"], e["is_on_trace"], e["too_close"], e["trail_len"]]
    names = ["step_count", "cur_target_id", "leader_finished", "finish_timer", "done", "crash", "is_in_box",
             "is_on_trace", "too_close", "trail_len"]
    for k, (g, n) in enumerate(zip(got_ints, names)):
        if t == 0 and n in ("is_in_box", "is_on_trace", "too_close"):
            continue
        _equal(int(g), int(ints[k]), n, t)
    if t > 0:
        _equal([int(e["mission_status"]), int(e["agent_status"]), int(e["leader_status"])],
               [int(x) for x in ints[10:13]], "info codes", t)
        _equal(int(out.done[env_index]), int(ints[4]), "out.done", t)
        _equal([int(x) for x in out.status[env_index]], [int(ints[10]), int(ints[11]), int(ints[12]), int(ints[5])],
               "out.status", t)
    rf, ri = d["t_robot_f"][t], d["t_robot_i"][t]
    for k, r in enumerate(_robots(e, nb)):
        nm = ("follower", "leader", "bear0", "bear1", "bear2", "bear3")[k]
        _equal([int(r["rot_dir"]), int(r["des_rot_dir"])] + [int(x) for x in r["rect"]], [int(x) for x in ri[k]],
               nm + " (rot_dir, des_rot_dir, rect)", t)
        got = [r["pos"][0], r["pos"][1], r["dir"], r["speed"], r["rot_speed"], r["des_speed"], r["des_rot_speed"]]
        if float_rtol == 0.0:
            _equal(np.array(got, np.float64), rf[k], nm + " floats (pos, dir, speed, rot_speed, des_*)", t)
        else:
            _close(got, rf[k], float_rtol, nm + " floats", t)
    if nb:
        _equal(e["bear_index"][:nb], d["t_bear_index"][t], "bear_index", t)
        if float_rtol == 0.0:
            _equal(e["bear_target"][:nb], d["t_bear_target"][t], "bear_target", t)
        else:
            _close(e["bear_target"][:nb], d["t_bear_target"][t], float_rtol, "bear_target", t)
    # ---- scalars --------------------------------------------------------------------------------
    fl = d["t_floats"][t]
    got = [e["last_reward"], e["overall_reward"], e["accumulated_penalty"], e["cur_speed_multiplier"]]
    if t == 0:
        got[0] = fl[0]
    if float_rtol == 0.0:
        _equal(np.array(got, np.float64), fl, "(reward, overall_reward, accumulated_penalty, speed_mult)", t)
    else:
        _close(got, fl, float_rtol, "(reward, overall_reward, accumulated_penalty, speed_mult)", t)
    if t > 0:
        _close(out.reward[env_index], fl[0], 1e-6, "out.reward", t)
    # ---- trail ----------------------------------------------------------------------------------
    tl = int(e["trail_len"])
    _equal(st.trail[env_index, tl - 1], d["t_trail_last"][t], "newest trail point", t)
    # ---- observation ----------------------------------------------------------------------------
    if float_rtol == 0.0:
        _equal(out.numerical_features[env_index], d["t_nf"][t], "numerical_features", t)
    else:
        _close(out.numerical_features[env_index], d["t_nf"][t], float_rtol, "numerical_features", t)
    _equal(out.leader_target[env_index], d["t_target"][t], "leader_target_point", t)
    # ---- tracker --------------------------------------------------------------------------------
    cap = c.corridor_cap
    if c.tracker_enabled:
        ti = d["t_tracker_i"][t]
        live = int(e["ring_head"]) - int(e["ring_tail"])
        n64 = max(0, min(int(e["hist_f64_end"]), int(e["ring_head"])) - int(e["ring_tail"]))
        _equal([int(e["saving_counter"]), live, live, n64], [int(x) for x in ti],
               "(saving_counter, len(hist), len(corridor), float64 points)", t)
        h, tail = int(e["ring_head"]), int(e["ring_tail"])
        _equal(st.hist[env_index, (h - 1) % cap], d["t_hist_last"][t], "hist[-1]", t)
        _equal(st.hist[env_index, tail % cap], d["t_hist_first"][t], "hist[0]", t)
        _equal(st.corridor[env_index, (h - 1) % cap], d["t_corr_last"][t].astype(np.float32), "corridor[-1] (f32)", t)
        _equal(st.corridor[env_index, tail % cap], d["t_corr_first"][t].astype(np.float32), "corridor[0] (f32)", t)
    # ---- FollowerInfo / LeaderTrackDetector_vector (SEN:834-842, 365-380) -------------------------
    if "t_follower_info" in d:
        if float_rtol == 0.0:
            _equal(out.follower_info[env_index], d["t_follower_info"][t], "FollowerInfo", t)
        else:
            _close(out.follower_info[env_index], d["t_follower_info"][t], float_rtol, "FollowerInfo", t)
    if "t_track_vectors" in d:
        want = d["t_track_vectors"][t]
        if float_rtol == 0.0:
            _equal(out.track_vectors[env_index], want, "LeaderTrackDetector_vector", t)
        else:   # differences of positions: absolute tolerance on the scale of the positions themselves
            err = np.abs(out.track_vectors[env_index].astype(np.float64) - want)
            if np.any(err > float_rtol * 1500.0):
                raise Mismatch("step %d: LeaderTrackDetector_vector differs by %g" % (t, err.max()))
    # ---- LeaderTrackDetector_radar (SEN:425-461) ------------------------------------------------------
    if "t_radar" in d:
        want = d["t_radar"][t]
        got = out.radar[env_index]
        if float_rtol == 0.0:
            _equal(got, want, "LeaderTrackDetector_radar", t)
        else:
            # a point on a sector boundary may land in the neighbouring sector (the boundary test is discontinuous):
            # such sectors must stay rare
            err = np.abs(got.astype(np.float64) - want) > float_rtol * np.maximum(1.0, np.abs(want))
            if err.sum() > max(2, 0.02 * err.size):
                raise Mismatch("step %d: LeaderTrackDetector_radar differs in %d of %d sectors" % (t, err.sum(), err.size))
    # ---- LaserSensor (SEN:63-136): sample points are discrete, so within tolerance means equal up to float rounding --
    if "t_laser" in d:
        want = d["t_laser"][t]
        got = out.laser[env_index]
        if float_rtol == 0.0:
            _equal(got, want, "LaserSensor", t)
        else:
            err = np.abs(got.astype(np.float64) - want) > float_rtol * np.maximum(1.0, np.abs(want))
            if err.sum() > max(2, 0.02 * err.size):   # a sample on a hit-box edge may fall either way
                raise Mismatch("step %d: LaserSensor differs in %d of %d values" % (t, err.sum(), err.size))
    # ---- rays -----------------------------------------------------------------------------------
    bad = 0
    if c.n_ray_sensors:
        got = np.asarray(out.rays[env_index, :gc.rays_per_env], np.float64)
        want = np.asarray(d["t_rays"][t], np.float64)
        if ray_rtol == 0.0:
            _equal(got.astype(np.float32), want.astype(np.float32), "rays", t)
        else:
            err = np.abs(got - want)
            tol = ray_rtol * np.maximum(1.0, np.abs(want))
            bad = int(np.sum(err > tol))
            if bad and ray_outlier_budget is None:
                i = int(np.argmax(err - tol))
                raise Mismatch("step %d: ray %d differs: got %r want %r" % (t, i, got[i], want[i]))
    return bad


def replay(stepper, d, gc, env_index=0, float_rtol=0.0, ray_rtol=RTOL, ray_outlier_budget=None, teacher=None,
           max_steps=None):
    """Reset the stepper on scenario 0 and replay the recorded actions, checking every record.
    All envs of the stepper receive the same action.  Returns (#steps checked, #ray outliers)."""
    n = stepper.n
    out = stepper.reset(scenario_ids=np.zeros(n, np.int32))
    st = stepper.get_state()
    outliers = check_record(0, d, gc, out, st, env_index, float_rtol, ray_rtol, ray_outlier_budget)
    acts = d["actions"]
    T = len(acts) if max_steps is None else min(len(acts), max_steps)
    has_inputs = "step_frames" in d
    for t in range(T):
        a = np.repeat(acts[t][None, :], n, axis=0)
        if has_inputs:   # what the reference drew from its global RNGs inside this step (FtlStepInputs)
            out = stepper.step(a, frames=np.full(n, d["step_frames"][t], np.int32),
                               regime_draws=np.repeat(d["step_draws"][t][None, :], n, axis=0))
        else:
            out = stepper.step(a)
        st = stepper.get_state()
        outliers += check_record(t + 1, d, gc, out, st, env_index, float_rtol, ray_rtol, ray_outlier_budget)
    if ray_outlier_budget is not None and outliers > ray_outlier_budget:
        raise Mismatch("%d ray values outside tolerance (budget %d)" % (outliers, ray_outlier_budget))
    return T, outliers


def check_final_arrays(st, d, gc, env_index=0):
    """Whole trail / tracker history / corridor at the end of the trace."""
    e = st.env[env_index]
    tl = int(e["trail_len"])
    _equal(st.trail[env_index, :tl], d["final_trail"], "whole trail", -1)
    if gc.c.tracker_enabled:
        cap = gc.c.corridor_cap
        idx = [(i % cap) for i in range(int(e["ring_tail"]), int(e["ring_head"]))]
        _equal(st.hist[env_index, idx], d["final_hist"], "whole tracker history", -1)
        _equal(st.corridor[env_index, idx], d["final_corridor"].astype(np.float32), "whole corridor (f32)", -1)


def sample_actions(gc, rng, n, t=0):
    """Uniform actions of the configured action space; every third step half of the batch drives straight ahead at
    full speed so that part of it stays near the trail (in-box rewards, finish timers)."""
    if gc.discrete_action_space:
        a = rng.randint(0, 5, size=n).astype(np.int32)
        if t % 3 == 0:
            a[: n // 2] = 2
        return a
    lo, hi = gc.action_bounds()
    a = rng.uniform(lo, hi, size=(n, len(lo))).astype(np.float32)
    if t % 3 == 0:
        a[: n // 2] = (0.0,) if len(lo) == 1 else (hi[0], 0.0)
    return a


def sensor_prev_expected(gc, rays):
    """ContinuousObserveModifier_sensorPrev.observation (WRP:203-221) applied with numpy to raw sensor blocks
    [N, rays_per_env]: clip(block / laser_length, 0, 1) per sensor, concatenated along the ray axis -> [N, H, sum W]."""
    feats = []
    for i, (_name, off, h, w) in enumerate(gc.ray_layout()):
        L = gc.c.ray[i].laser_length
        block = rays[:, off:off + h * w].reshape(rays.shape[0], h, w)   # float32 / python float -> float32 (NEP 50)
        feats.append(np.clip(block / L, 0, 1))
    return np.concatenate(feats, axis=2)
