"""ctypes mirror of include/ftl.h (the C-ABI of libftl.so).

Every structure here must stay field-for-field identical to the header; tests/test_abi.py checks
the sizes against the compiled libraries.
"""
import ctypes as C

import numpy as np

FTL_ABI_VERSION = 4
FTL_MAX_BEARS = 4
FTL_MAX_RAY_SENSORS = 4
FTL_MAX_REGIME = 16
FTL_MAX_HIST = 8
FTL_MAX_CUSTOM_ANGLES = 8

FTL_OK = 0
FTL_ERR_INVALID, FTL_ERR_CUDA, FTL_ERR_STATE, FTL_ERR_NOMEM = -1, -2, -3, -4

MISSION_STATUS = ("in_progress", "fail", "success", "finished_by_time")
AGENT_STATUS = ("moving", "crash", "finished", "low_reward", "too_far_from_leader")
LEADER_STATUS = ("moving", "finished", "crash")

REACT_NONE, REACT_ALL, REACT_STATIC, REACT_DYNAMIC = 0, 1, 2, 3
ACTION_CONTINUOUS, ACTION_CONST_SPEED, ACTION_DISCRETE = 0, 1, 2

(STAT_EPISODES, STAT_RETURN_SUM, STAT_LENGTH_SUM, STAT_CRASH, STAT_SUCCESS, STAT_TIMEOUT,
 STAT_LEADER_CRASH, STAT_ENV_STEPS, STAT_OVERFLOW) = range(9)
STAT_COUNT = 16
STAT_NAMES = ("episodes", "return_sum", "length_sum", "crash", "success", "timeout", "leader_crash",
              "env_steps", "overflow")


class FtlRobotConfig(C.Structure):
    _fields_ = [("min_speed", C.c_double), ("max_speed", C.c_double), ("max_rotation_speed", C.c_double),
                ("max_speed_change", C.c_double), ("max_rotation_speed_change", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32)]


class FtlRaySensorConfig(C.Structure):
    _fields_ = [("lasers_count", C.c_int32), ("max_prev_obs", C.c_int32), ("pad_sectors", C.c_int32),
                ("react_to_safe_corridor", C.c_int32), ("react_to_green_zone", C.c_int32),
                ("react_to_obstacles", C.c_int32), ("laser_length", C.c_double),
                ("first_laser_angle_offset", C.c_double),
                ("n_custom_angles", C.c_int32), ("compas", C.c_int32),
                ("custom_angle", C.c_double * FTL_MAX_CUSTOM_ANGLES)]


class FtlConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("game_width", C.c_int32), ("game_height", C.c_int32),
        ("frames_per_step", C.c_int32), ("max_steps", C.c_int32), ("warm_start", C.c_int32),
        ("trajectory_saving_period", C.c_int32), ("aggregate_reward", C.c_int32),
        ("ignore_follower_collisions", C.c_int32), ("action_mode", C.c_int32),
        ("leader_pos_epsilon", C.c_double),
        ("min_distance", C.c_double), ("max_distance", C.c_double), ("max_dev", C.c_double),
        ("const_speed_action", C.c_double), ("discrete_rotation_table", C.c_double * 5),
        ("follower", FtlRobotConfig), ("leader", FtlRobotConfig), ("bear", FtlRobotConfig),
        ("n_bears", C.c_int32), ("move_bear_v4", C.c_int32),
        ("reward_in_box", C.c_double), ("reward_on_track", C.c_double), ("reward_in_dev", C.c_double),
        ("leader_movement_reward", C.c_double), ("crash_penalty", C.c_double),
        ("not_on_track_penalty", C.c_double), ("too_close_penalty", C.c_double),
        ("leader_stop_penalty", C.c_double),
        ("es_has_low_reward", C.c_int32), ("es_has_max_distance_coef", C.c_int32),
        ("es_low_reward", C.c_double), ("es_max_distance_coef", C.c_double),
        ("n_speed_regime", C.c_int32),
        ("speed_regime_key", C.c_int32 * FTL_MAX_REGIME),
        ("speed_regime_is_range", C.c_int32 * FTL_MAX_REGIME),
        ("speed_regime_lo", C.c_double * FTL_MAX_REGIME), ("speed_regime_hi", C.c_double * FTL_MAX_REGIME),
        ("n_accel_regime", C.c_int32),
        ("accel_regime_key", C.c_int32 * FTL_MAX_REGIME),
        ("accel_regime_val", C.c_double * FTL_MAX_REGIME),
        ("tracker_enabled", C.c_int32), ("saving_period", C.c_int32),
        ("start_corridor_behind_follower", C.c_int32), ("tracker_scans_per_step", C.c_int32),
        ("corridor_length", C.c_double), ("corridor_width", C.c_double),
        ("n_ray_sensors", C.c_int32),
        ("ray", FtlRaySensorConfig * FTL_MAX_RAY_SENSORS),
        ("trail_cap", C.c_int32), ("corridor_cap", C.c_int32), ("route_cap", C.c_int32),
        ("static_cap", C.c_int32), ("auto_reset", C.c_int32),
        ("fused_sensor_prev", C.c_int32),
        ("track_vector_len", C.c_int32), ("track_vector_mode", C.c_int32),
        ("radar_sectors", C.c_int32), ("radar_len", C.c_int32), ("radar_mode", C.c_int32),
        ("laser_points", C.c_int32), ("laser_beams", C.c_int32), ("laser_only_distances", C.c_int32),
        ("laser_available_angle", C.c_double), ("laser_angle_step", C.c_double),
        ("laser_range", C.c_double), ("laser_reach_extra", C.c_double),
    ]


class FtlScenarioPool(C.Structure):
    _fields_ = [
        ("n_scenarios", C.c_int32), ("static_cap", C.c_int32), ("route_cap", C.c_int32),
        ("static_rects", C.c_void_p), ("n_static", C.c_void_p),
        ("route", C.c_void_p), ("n_route", C.c_void_p),
        ("leader_pos", C.c_void_p), ("leader_dir", C.c_void_p),
        ("follower_pos", C.c_void_p), ("follower_dir", C.c_void_p),
        ("found_target_point", C.c_void_p),
    ]


class FtlScenarioGenConfig(C.Structure):
    _fields_ = [("game_width", C.c_int32), ("game_height", C.c_int32),
                ("min_distance", C.c_double), ("max_distance", C.c_double), ("leader_pos_epsilon", C.c_double),
                ("leader_width", C.c_int32), ("leader_height", C.c_int32),
                ("follower_width", C.c_int32), ("follower_height", C.c_int32),
                ("leader_width_f", C.c_double), ("leader_height_f", C.c_double),
                ("add_obstacles", C.c_int32), ("obstacle_number", C.c_int32), ("step_grid", C.c_int32),
                ("bridge_size", C.c_int32 * 2), ("leader_margin", C.c_double),
                ("path_finding", C.c_int32), ("multiple_end_points", C.c_int32)]


class FtlRobotState(C.Structure):
    _fields_ = [("pos", C.c_float * 2), ("rect", C.c_int32 * 4), ("dir", C.c_double),
                ("speed", C.c_double), ("rot_speed", C.c_double),
                ("des_speed", C.c_double), ("des_rot_speed", C.c_double),
                ("rot_dir", C.c_int32), ("des_rot_dir", C.c_int32)]


class FtlSnapshot(C.Structure):
    _fields_ = [("valid", C.c_int32), ("corr_tail", C.c_int32), ("corr_head", C.c_int32), ("pad_", C.c_int32),
                ("dyn_rect", (C.c_int32 * 4) * (1 + FTL_MAX_BEARS))]


class FtlEnvState(C.Structure):
    _fields_ = [
        ("follower", FtlRobotState), ("leader", FtlRobotState), ("bear", FtlRobotState * FTL_MAX_BEARS),
        ("bear_target", (C.c_double * 2) * FTL_MAX_BEARS),
        ("bear_index", C.c_int32 * FTL_MAX_BEARS),
        ("accumulated_penalty", C.c_double), ("overall_reward", C.c_double), ("last_reward", C.c_double),
        ("cur_speed_multiplier", C.c_double), ("cur_leader_acceleration", C.c_double),
        ("cur_leader_cumulative_speed", C.c_double),
        ("accel_consumed", C.c_int32), ("scenario_id", C.c_int32),
        ("cur_target_id", C.c_int32), ("leader_finished", C.c_int32),
        ("step_count", C.c_int32), ("finish_timer", C.c_int32),
        ("done", C.c_int32), ("crash", C.c_int32), ("is_in_box", C.c_int32), ("is_on_trace", C.c_int32),
        ("too_close", C.c_int32),
        ("mission_status", C.c_int32), ("agent_status", C.c_int32), ("leader_status", C.c_int32),
        ("trail_len", C.c_int32), ("saving_counter", C.c_int32),
        ("ring_tail", C.c_int32), ("ring_head", C.c_int32), ("hist_f64_end", C.c_int32),
        ("snap_pushes", C.c_int32), ("episode_count", C.c_int32), ("overflow", C.c_int32), ("pad_", C.c_int32),
        ("snap", FtlSnapshot * FTL_MAX_HIST),
    ]


class FtlStateBuffers(C.Structure):
    _fields_ = [("env", C.c_void_p), ("trail", C.c_void_p), ("hist", C.c_void_p), ("corridor", C.c_void_p)]


class FtlOutputs(C.Structure):
    _fields_ = [("numerical_features", C.c_void_p), ("leader_target", C.c_void_p), ("rays", C.c_void_p),
                ("reward", C.c_void_p), ("done", C.c_void_p), ("status", C.c_void_p),
                ("follower_info", C.c_void_p), ("track_vectors", C.c_void_p), ("radar", C.c_void_p),
                ("laser", C.c_void_p)]


class FtlStepInputs(C.Structure):
    """Optional per-step inputs of ftl_step_ex / ftl_step_host_ex (include/ftl.h)."""
    _fields_ = [("frames_per_step", C.c_void_p), ("regime_draws", C.c_void_p)]


class FtlMlpWeights(C.Structure):
    """Device pointers of the fused policy kernel's weights (ftl_policy_mlp, include/ftl.h)."""
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p),
                ("b3", C.c_void_p), ("noise_scale", C.c_void_p), ("act_mid", C.c_void_p), ("act_half", C.c_void_p),
                ("obs_dim", C.c_int32), ("act_dim", C.c_int32)]


OPT_KIN_PDL, OPT_STEP_PHASE, OPT_NO_OVERLAP = 1, 2, 3   # ftl_set_option

ENV_STATE_DTYPE = np.dtype(FtlEnvState)


def laser_beam_count(available_angle, angle_step):
    """Beams of a LaserSensor: -direction, then +-k * angle_step while the running offset is below int(angle / 2), SEN:86-98."""
    border, diff, n = int(min(360, available_angle) / 2), 0, 1
    while diff < border:
        diff += angle_step
        n += 2
    return n


def rays_per_env(cfg):
    n = 0
    for s in range(cfg.n_ray_sensors):
        sc = cfg.ray[s]
        n += sc.max_prev_obs * (5 * sc.lasers_count if sc.compas else 4 * sc.lasers_count if sc.pad_sectors
                                else sc.lasers_count)
    return n


def ptr(a):
    """Host pointer of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
