"""Constructor arguments of the reference ``Game`` -> the POD ``FtlConfig`` the kernels consume.

Keeps the keyword names and defaults of ``Game.__init__`` verbatim
(src/continuous_grid_arctic/follow_the_leader_continuous_env.py:45-105), does the same unit
conversions (metres -> pixels ENV:283-285, m/s -> px/frame with AVG_FRAMES_PER_SECOND=100
ENV:38,330-357, bears ENV:704-714) and raises the same exception types for the same mistakes
(ENV:241, 419-427; CLS:242-249; SEN:761, 847-851).
"""
import math
from collections import OrderedDict
from warnings import warn

import numpy as np

from . import abi

AVG_FRAMES_PER_SECOND = 100  # ENV:38

REWARD_DEFAULTS = dict(  # utils/reward_constructor.py:4-15
    name="base_reward", reward_in_box=1., reward_on_track=0.1, reward_in_dev=0.5, leader_movement_reward=1.,
    crash_penalty=-10., not_on_track_penalty=-1., too_close_penalty=-5., leader_stop_penalty=-1.)

GAME_DEFAULTS = OrderedDict(  # ENV:45-105
    game_width=1500, game_height=1000, framerate=500, frames_per_step=10, random_frames_per_step=None,
    caption="Serious Robot Follower Simulation v.-1", trajectory=None, leader_pos_epsilon=25,
    show_leader_path_flag=True, show_leader_trajectory_flag=True, show_rectangles_flag=True,
    show_box_flag=True, show_objects_flag=True, show_sensors_flag=True, simulation_time_limit=None,
    reward_config=None, pixels_to_meter=50, min_distance=1, max_distance=4, max_dev=1, warm_start=500,
    manual_control=False, manual_control_input="keyboard", max_steps=5000, aggregate_reward=False,
    add_obstacles=True, add_bear=True, bear_number=3, multi_random_bears=False, move_bear_v4=True,
    obstacle_number=35, bear_behind=False, step_grid=10, early_stopping={}, follower_sensors={},
    leader_speed_regime=None, leader_acceleration_regime=None, discrete_action_space=False,
    constant_follower_speed=False, path_finding_algorythm="dstar", multiple_end_points=False,
    negative_speed=False, follower_max_speed=0.5, leader_max_speed=0.5,
    follower_max_rotation_speed=57.296, leader_max_rotation_speed=57.296, follower_acceleration=0.005,
    leader_acceleration=0.005, bear_max_speed=1.1, follower_size=(0.5, 0.35), leader_size=(0.38, 0.52),
    bear_size=(0.5, 0.5), bridge_size=(80, 40), return_render_matrix=True,
    ignore_follower_collisions=False, path_finding_iterations=15000, leader_margin=1.5)

TRACKER_CLASSES = ("LeaderPositionsTracker_v2",)
RAY_CLASSES = ("LeaderCorridor_Prev_lasers_v2", "LaserPrevSensor", "LeaderCorridor_lasers_compas")
# sensors without history on the same ray engine (max_prev_obs = 1, output of shape (R,)): SEN:571-807
FLAT_RAY_CLASSES = ("LeaderCorridor_lasers", "LeaderCorridor_lasers_v2")
# Every class name the reference registry knows (SEN:1291-1307); anything else is "undefined".
KNOWN_SENSOR_CLASSES = (
    "LaserSensor", "LeaderPositionsTracker", "LeaderPositionsTracker_v2", "LeaderTrackDetector_vector",
    "LeaderTrackDetector_radar", "LeaderCorridor_lasers", "GreenBoxBorderSensor", "LeaderCorridor_lasers_v2",
    "LeaderObstacles_lasers", "Leader_Dyn_Obstacles_lasers", "FollowerInfo", "LaserPrevSensor",
    "LeaderCorridor_Prev_lasers_v2", "LeaderCorridor_Prev_lasers_v3", "LeaderCorridor_lasers_compas")

_REACT = {False: abi.REACT_NONE, True: abi.REACT_ALL, "all": abi.REACT_ALL, "static": abi.REACT_STATIC,
          "dynamic": abi.REACT_DYNAMIC}


def _pow2_at_least(n):
    p = 1
    while p < n:
        p *= 2
    return p


class GameConfig:
    """Parsed constructor arguments; ``.c`` is the ``FtlConfig`` struct handed to libftl."""

    def __init__(self, trail_cap=None, corridor_cap=None, route_cap=None, static_cap=None, auto_reset=False,
                 strict_lasers_count=False, fused_sensor_prev=False, **kwargs):
        kw = OrderedDict(GAME_DEFAULTS)
        extra = {k: v for k, v in kwargs.items() if k not in kw}
        kw.update({k: v for k, v in kwargs.items() if k in kw})
        self.kwargs = kw
        self.extra_kwargs = extra  # Game.__init__ swallows unknown names through **kwargs (ENV:104)
        g = kw
        self.strict_lasers_count = strict_lasers_count

        # ---- validation with the reference's exception types ---------------------------------
        if g["multiple_end_points"] and g["path_finding_algorythm"] != "dstar":  # ENV:239-243
            raise NotImplementedError("Only dstar pathfinding function supports multiple end points. "
                                      "multiple_end_points must be False or diggerent path_finding_algorythm "
                                      "must be chosen")
        if g["path_finding_algorythm"] not in ["astar", "dstar"]:  # ENV:423-425
            raise ValueError("path_finding_algorythm {} not in list:{}".format(
                g["path_finding_algorythm"], ["astar", "dstar"]))
        if g["add_bear"] and g["bear_number"] <= 0:  # ENV:426-427
            raise ValueError("Add bear is true, but number of bears is not greater then 0")
        if g["manual_control"]:
            raise NotImplementedError("manual_control needs a pygame window; the batched simulator is head-less")
        if g["simulation_time_limit"] is not None:
            raise NotImplementedError("simulation_time_limit is wall-clock based (ENV:1119-1123) and not supported")
        # random_frames_per_step (ENV:398-405, 939-940): the reference redraws self.frames_per_step with
        # np.random.randint(lo, hi) after every step.  Here the kernels take the per-env count of each step as an input
        # (FtlStepInputs.frames_per_step) and FtlConfig.frames_per_step becomes the capacity hi - 1.
        self.random_frames_per_step = None
        if g["random_frames_per_step"] is not None:
            rf = g["random_frames_per_step"]
            if g["frames_per_step"] is not None:
                warn("random_frames_per_step and frames_per_step are both given; random_frames_per_step is used")
            assert len(rf) == 2, ("random_frames_per_step must be the two bounds of the random draw. "
                                  "Given: {}".format(rf))
            lo_f, hi_f = int(rf[0]), int(rf[1])
            if lo_f < 1 or hi_f <= lo_f:
                raise ValueError("random_frames_per_step needs 1 <= low < high (np.random.randint(low, high))")
            self.random_frames_per_step = (lo_f, hi_f)

        ptm = g["pixels_to_meter"]
        px = lambda m: m * ptm  # noqa: E731  ENV:1942-1943

        c = abi.FtlConfig()
        c.abi_version = abi.FTL_ABI_VERSION
        # library option: `rays` carries ContinuousObserveModifier_sensorPrev's output (WRP:203-221) directly
        c.fused_sensor_prev = int(bool(fused_sensor_prev))
        c.game_width, c.game_height = int(g["game_width"]), int(g["game_height"])
        c.frames_per_step = self.random_frames_per_step[1] - 1 if self.random_frames_per_step else int(g["frames_per_step"])
        c.max_steps, c.warm_start = int(g["max_steps"]), int(g["warm_start"])
        c.trajectory_saving_period = 5  # ENV:262
        c.aggregate_reward = int(bool(g["aggregate_reward"]))
        c.ignore_follower_collisions = int(bool(g["ignore_follower_collisions"]))
        c.leader_pos_epsilon = float(g["leader_pos_epsilon"])
        c.min_distance, c.max_distance, c.max_dev = float(px(g["min_distance"])), float(px(g["max_distance"])), \
            float(px(g["max_dev"]))

        # ---- robots, ENV:330-357 -----------------------------------------------------------------
        fmax = px(g["follower_max_speed"]) / AVG_FRAMES_PER_SECOND
        self._robot(c.follower, min_speed=-fmax if g["negative_speed"] else 0, max_speed=fmax,
                    max_rot=g["follower_max_rotation_speed"] / AVG_FRAMES_PER_SECOND,
                    accel=px(g["follower_acceleration"]) / AVG_FRAMES_PER_SECOND,
                    height=px(g["follower_size"][0]), width=px(g["follower_size"][1]))
        lmax = px(g["leader_max_speed"]) / AVG_FRAMES_PER_SECOND
        self._robot(c.leader, min_speed=0, max_speed=lmax,
                    max_rot=g["leader_max_rotation_speed"] / AVG_FRAMES_PER_SECOND,
                    accel=px(g["leader_acceleration"]) / AVG_FRAMES_PER_SECOND,
                    width=px(g["leader_size"][0]), height=px(g["leader_size"][1]))
        self._robot(c.bear, min_speed=0, max_speed=g["bear_max_speed"] * lmax,  # ENV:704-714
                    max_rot=g["leader_max_rotation_speed"] / AVG_FRAMES_PER_SECOND, accel=px(0.005),
                    height=px(g["bear_size"][0]), width=px(g["bear_size"][1]))
        c.n_bears = int(g["bear_number"]) if g["add_bear"] else 0
        if c.n_bears > abi.FTL_MAX_BEARS:
            raise ValueError("at most %d dynamic obstacles are supported (bear indices > 3 draw random "
                             "targets in the reference, ENV:750-754)" % abi.FTL_MAX_BEARS)
        c.move_bear_v4 = int(bool(g["move_bear_v4"]))

        # ---- action space, ENV:360-378 -----------------------------------------------------------
        self.discrete_action_space = bool(g["discrete_action_space"])
        self.constant_follower_speed = bool(g["constant_follower_speed"])
        mr = c.follower.max_rotation_speed
        for i, v in enumerate((-mr, -mr / 2, 0, mr / 2, mr)):
            c.discrete_rotation_table[i] = v
        c.const_speed_action = 0.25  # ENV:925
        if self.discrete_action_space:
            c.action_mode = abi.ACTION_DISCRETE
        elif self.constant_follower_speed:
            c.action_mode = abi.ACTION_CONST_SPEED
        else:
            c.action_mode = abi.ACTION_CONTINUOUS

        # ---- reward, ENV:276-279 -----------------------------------------------------------------
        rw = dict(REWARD_DEFAULTS)
        if g["reward_config"]:
            import json
            with open(g["reward_config"], "r") as f:
                rw.update(json.load(f))
        else:
            rw["leader_movement_reward"] = 0
        self.reward = rw
        for k in ("reward_in_box", "reward_on_track", "reward_in_dev", "leader_movement_reward", "crash_penalty",
                  "not_on_track_penalty", "too_close_penalty", "leader_stop_penalty"):
            setattr(c, k, float(rw[k]))

        # ---- early stopping, ENV:1088-1107 -------------------------------------------------------
        es = g["early_stopping"] or {}
        c.es_has_low_reward = int("low_reward" in es)
        c.es_low_reward = float(es.get("low_reward", 0.0))
        c.es_has_max_distance_coef = int("max_distance_coef" in es)
        c.es_max_distance_coef = float(es.get("max_distance_coef", 0.0))

        # ---- leader regimes, ENV:381-397 ---------------------------------------------------------
        self._regimes(c, g["leader_speed_regime"], g["leader_acceleration_regime"])

        # ---- sensors -----------------------------------------------------------------------------
        self.follower_sensors = g["follower_sensors"]
        self.ray_sensor_names = []
        self._sensors(c, g["follower_sensors"])

        # ---- capacities --------------------------------------------------------------------------
        lead_v = max(c.leader.max_speed, 1e-9)
        init_trail = int(px(g["max_distance"]) * 0.9 / (5 * lead_v)) + 2  # ENV:535-536 with the 0.9*max start gap
        want_trail = init_trail + c.max_steps // 5 + 8
        c.trail_cap = int(trail_cap) if trail_cap else int(math.ceil(want_trail / 32.0) * 32)
        if corridor_cap is None:
            corridor_cap = 64
            if c.tracker_enabled:
                f_min = self.random_frames_per_step[0] if self.random_frames_per_step else c.frames_per_step
                per_save = max(lead_v * f_min * c.saving_period / max(c.tracker_scans_per_step, 1), 1e-6)
                corridor_cap = min(_pow2_at_least(int(1.5 * c.corridor_length / per_save) + 32), 512)
        if corridor_cap & (corridor_cap - 1) or corridor_cap > 512:
            raise ValueError("corridor_cap must be a power of two <= 512")
        c.corridor_cap = int(corridor_cap)
        # three D* legs across the field (multiple_end_points) are up to ~4 x a single route
        c.route_cap = int(route_cap) if route_cap else (512 if g["multiple_end_points"] else 128)
        c.static_cap = int(static_cap) if static_cap else max(2 + int(g["obstacle_number"]) if g["add_obstacles"] else 0, 1)
        c.auto_reset = int(bool(auto_reset))
        self.c = c

    # -------------------------------------------------------------------------------------------
    @staticmethod
    def _robot(rc, min_speed, max_speed, max_rot, accel, width, height):
        rc.min_speed, rc.max_speed = float(min_speed), float(max_speed)
        rc.max_rotation_speed = float(max_rot)
        rc.max_speed_change = float(accel)
        rc.max_rotation_speed_change = 20 / 100  # ENV:564, 587, 712
        rc.width, rc.height = int(width), int(height)  # transform.scale truncates, CLS:42

    @staticmethod
    def _regimes(c, speed, accel):
        c.n_speed_regime = 0
        if type(speed) in (dict, OrderedDict):
            if len(speed) > abi.FTL_MAX_REGIME:
                raise ValueError("leader_speed_regime: at most %d keys" % abi.FTL_MAX_REGIME)
            for i, (k, v) in enumerate(speed.items()):
                c.speed_regime_key[i] = int(k)
                if type(v) in (tuple, list):
                    c.speed_regime_is_range[i] = 1
                    c.speed_regime_lo[i], c.speed_regime_hi[i] = float(v[0]), float(v[1])
                else:
                    c.speed_regime_is_range[i] = 0
                    c.speed_regime_lo[i] = c.speed_regime_hi[i] = float(v)
            c.n_speed_regime = len(speed)
        elif speed is not None:
            warn("leader_speed_regime must be dict or OrderedDict, got {}; ignored".format(type(speed)))
        c.n_accel_regime = 0
        if type(accel) in (dict, OrderedDict):
            if len(accel) > abi.FTL_MAX_REGIME:
                raise ValueError("leader_acceleration_regime: at most %d keys" % abi.FTL_MAX_REGIME)
            for i, (k, v) in enumerate(accel.items()):
                c.accel_regime_key[i] = int(k)
                c.accel_regime_val[i] = float(v)
            c.n_accel_regime = len(accel)
        elif accel is not None:
            warn("leader_acceleration_regime must be dict, got {}; ignored".format(type(accel)))

    def _sensors(self, c, sensors):
        c.tracker_enabled = 0
        c.tracker_scans_per_step = 2  # CLS:263-286: the v2 tracker is scanned twice per use_sensors
        c.n_ray_sensors = 0
        c.track_vector_len = 0
        c.radar_sectors = 0
        self.radar_name = None             # dict key of the LeaderTrackDetector_radar sensor, if any
        self.laser_name = None             # dict key of the LaserSensor, if any
        self.laser_shape = None            # per-env shape of its output: (beams, 2) or (beams,)
        c.laser_points = 0
        self.ray_sensor_flat = []          # True: the sensor returns (R,), not (H, R)
        self.follower_info_name = None     # dict key of the FollowerInfo sensor, if any
        self.track_vector_name = None      # dict key of the LeaderTrackDetector_vector sensor, if any
        for k in (sensors or {}):          # CLS:240-244
            if k in ("LeaderTrackDetector_vector", "LeaderTrackDetector_radar") and \
                    "LeaderPositionsTracker" not in sensors and "LeaderPositionsTracker_v2" not in sensors:
                raise ValueError(
                    "Sensor {} requires sensor LeaderPositionsTracker for tracking leader movement.".format(k))
        for name, sc in (sensors or {}).items():
            cls = sc.get("sensor_class", name)
            if cls not in KNOWN_SENSOR_CLASSES:  # CLS:244-249
                raise ValueError(f"Sensor class is undefined: {name}")
            args = {k: v for k, v in sc.items() if k != "sensor_class"}
            if cls in TRACKER_CLASSES:
                if name != "LeaderPositionsTracker_v2":
                    # use_sensors only looks the tracker up under these literal keys (CLS:257, 263);
                    # under any other key the reference dies with an unbound leader_corridor.
                    raise ValueError("the tracker must be registered under the key 'LeaderPositionsTracker_v2'")
                # eat_close_points is accepted and ignored, as upstream: only the deprecated LeaderPositionsTracker reads
                # it (SEN:207); LeaderPositionsTracker_v2.scan (SEN:243-327) never does
                if not args.get("generate_corridor", True):
                    # upstream this configuration dies with IndexError at the first trim of the history: scan pops the
                    # (empty) corridor deque together with the history (SEN:292-293)
                    raise NotImplementedError("generate_corridor=False is not supported (the reference itself raises "
                                              "IndexError once the history exceeds corridor_length, SEN:292-293)")
                c.tracker_enabled = 1
                c.saving_period = int(args.get("saving_period", 5))
                c.start_corridor_behind_follower = int(bool(args.get("start_corridor_behind_follower", False)))
                c.corridor_length = float(args["corridor_length"])  # required keyword, SEN:237
                c.corridor_width = float(args["corridor_width"])
            elif cls in RAY_CLASSES:
                if c.n_ray_sensors >= abi.FTL_MAX_RAY_SENSORS:
                    raise ValueError("at most %d ray sensors" % abi.FTL_MAX_RAY_SENSORS)
                r = c.ray[c.n_ray_sensors]
                legacy = cls == "LaserPrevSensor"  # retired name -> same class with these flags (SEN:847-851)
                r.lasers_count = int(args.get("lasers_count", 12))
                if self.strict_lasers_count and r.lasers_count not in [12, 24, 20, 36]:  # SEN:761-762
                    raise ValueError("Invalid number of laser beams, should be 12,24,20 or 36")
                if r.lasers_count < 1:
                    raise ValueError("lasers_count must be positive")
                r.laser_length = float(args.get("laser_length", 100))
                r.max_prev_obs = int(args.get("max_prev_obs", 0))
                assert r.max_prev_obs > 0  # SEN:876
                if r.max_prev_obs > abi.FTL_MAX_HIST:
                    raise ValueError("max_prev_obs > %d is not supported" % abi.FTL_MAX_HIST)
                r.pad_sectors = int(bool(args.get("pad_sectors", True)))
                r.react_to_safe_corridor = int(bool(args.get("react_to_safe_corridor", not legacy)))
                r.react_to_green_zone = int(bool(args.get("react_to_green_zone", False)))
                rto = args.get("react_to_obstacles", True if legacy else False)
                if rto not in _REACT:
                    raise ValueError("You need to specify which obstacles the sensor should respond to. Set "
                                     "react_to_obstacles equal to one of the values: True, 'all', 'dynamic', 'static'")
                r.react_to_obstacles = _REACT[rto]
                r.first_laser_angle_offset = float(args.get("first_laser_angle_offset", 0 if legacy else -45))
                r.n_custom_angles = 0
                r.compas = int(cls == "LeaderCorridor_lasers_compas")
                if r.compas and (not r.react_to_safe_corridor or not r.react_to_green_zone or
                                 r.react_to_obstacles != abi.REACT_NONE):   # SEN:1148-1152
                    raise ValueError("Unsupported set of flags for LeaderCorridor_lasers_compas class, now implemented"
                                     "only option for flags: "
                                     "react_to_safe_corridor=True, react_to_green_zone=True, react_to_obstacles=False")
                self.ray_sensor_names.append(name)
                self.ray_sensor_flat.append(False)
                c.n_ray_sensors += 1
            elif cls in FLAT_RAY_CLASSES:
                if c.n_ray_sensors >= abi.FTL_MAX_RAY_SENSORS:
                    raise ValueError("at most %d ray sensors" % abi.FTL_MAX_RAY_SENSORS)
                r = c.ray[c.n_ray_sensors]
                if cls == "LeaderCorridor_lasers":   # SEN:577-606, 675-700: a fixed fan
                    front, back = int(args.get("front_lasers_count", 3)), int(args.get("back_lasers_count", 0))
                    assert front in [3, 5]           # SEN:597-598
                    assert back in [0, 2]
                    angles = [-40.0, 0.0, 40.0] + ([-90.0, 90.0] if front == 5 else []) + \
                             ([-150.0, 150.0] if back == 2 else [])
                    r.lasers_count = r.n_custom_angles = len(angles)
                    for k, a in enumerate(angles):
                        r.custom_angle[k] = a
                else:                                # SEN:742-769
                    r.lasers_count = int(args.get("lasers_count", 12))
                    if self.strict_lasers_count and r.lasers_count not in [12, 24, 20, 36]:
                        raise ValueError("Invalid number of laser beams, should be 12,24,20 or 36")
                    if r.lasers_count < 1:
                        raise ValueError("lasers_count must be positive")
                    r.n_custom_angles = 0
                r.laser_length = float(args.get("laser_length", 100))
                r.max_prev_obs, r.pad_sectors = 1, 0
                r.react_to_safe_corridor = int(bool(args.get("react_to_safe_corridor", True)))
                r.react_to_green_zone = int(bool(args.get("react_to_green_zone", False)))
                rto = args.get("react_to_obstacles", False)
                if rto not in _REACT:
                    raise ValueError("You need to specify which obstacles the sensor should respond to. Set "
                                     "react_to_obstacles equal to one of the values: True, 'all', 'dynamic', 'static'")
                r.react_to_obstacles = _REACT[rto]
                r.first_laser_angle_offset = 0.0
                self.ray_sensor_names.append(name)
                self.ray_sensor_flat.append(True)
                c.n_ray_sensors += 1
            elif cls == "FollowerInfo":              # SEN:822-842
                if int(args.get("speed_direction_param", 2)) != 2:
                    raise NotImplementedError("FollowerInfo: only speed_direction_param=2 (the default) is supported")
                if self.follower_info_name is not None:
                    raise NotImplementedError("one FollowerInfo sensor at most")
                self.follower_info_name = name
            elif cls == "LeaderTrackDetector_vector":  # SEN:349-380
                if self.track_vector_name is not None:
                    raise NotImplementedError("one LeaderTrackDetector_vector sensor at most")
                mode = args.get("detectable_positions", "new")
                if mode not in ("new", "old"):
                    raise ValueError("detectable_positions must be 'new' or 'old'")
                c.track_vector_len = int(args.get("position_sequence_length", 100))
                if c.track_vector_len < 1:
                    raise ValueError("position_sequence_length must be positive")
                c.track_vector_mode = 0 if mode == "new" else 1
                self.track_vector_name = name
            elif cls == "LeaderTrackDetector_radar":   # SEN:394-461
                if self.radar_name is not None:
                    raise NotImplementedError("one LeaderTrackDetector_radar sensor at most")
                mode = args.get("detectable_positions", "old")
                if mode not in ("new", "old", "near"):
                    raise ValueError("detectable_positions must be 'new', 'old' or 'near'")
                c.radar_len = int(args.get("position_sequence_length", 100))
                c.radar_sectors = int(args.get("radar_sectors_number", 180))
                if c.radar_len < 1 or c.radar_sectors < 1:
                    raise ValueError("position_sequence_length and radar_sectors_number must be positive")
                c.radar_mode = {"new": 0, "old": 1, "near": 2}[mode]
                self.radar_name = name
            elif cls == "LaserSensor":                 # SEN:18-136
                if self.laser_name is not None:
                    raise NotImplementedError("one LaserSensor at most")
                if args.get("return_all_points", False):
                    raise NotImplementedError("LaserSensor: return_all_points=True is not supported")
                c.laser_available_angle = float(min(360, args.get("available_angle", 360)))
                c.laser_angle_step = float(args.get("angle_step", 10))
                if c.laser_angle_step <= 0:
                    raise ValueError("LaserSensor: angle_step must be positive")
                c.laser_points = int(args.get("points_number", 20))
                if c.laser_points < 1:
                    raise ValueError("LaserSensor: points_number must be positive")
                c.laser_range = float(args.get("sensor_range", 5) * self.kwargs["pixels_to_meter"])
                c.laser_reach_extra = float(3 * self.kwargs["pixels_to_meter"])
                c.laser_only_distances = int(bool(args.get("return_only_distances", False)))
                c.laser_beams = abi.laser_beam_count(c.laser_available_angle, c.laser_angle_step)
                self.laser_name = name
                self.laser_shape = (c.laser_beams,) if c.laser_only_distances else (c.laser_beams, 2)
            else:
                raise NotImplementedError(
                    "sensor class %s is outside the accelerated path (SURVEY.md section 8(f)3)" % cls)
        if (c.n_ray_sensors or c.track_vector_len or c.radar_sectors) and not c.tracker_enabled:
            raise ValueError("ray sensors and track detectors need the LeaderPositionsTracker_v2 corridor (CLS:263-280)")

    # ---- spaces (ENV:360-378, 1812-1824) -----------------------------------------------------------
    def action_bounds(self):
        f = self.c.follower
        if self.discrete_action_space:
            return None
        if self.constant_follower_speed:
            return (np.array([-f.max_rotation_speed], np.float32), np.array([f.max_rotation_speed], np.float32))
        return (np.array((f.min_speed, -f.max_rotation_speed), dtype=np.float32),
                np.array((f.max_speed, f.max_rotation_speed), dtype=np.float32))

    def observation_bounds(self):
        c = self.c
        low = np.array((0, 0, 0, 0, -c.leader.max_rotation_speed, 0, 0, 0, 0, -c.follower.max_rotation_speed),
                       dtype=np.float32)
        high = np.array((c.game_width, c.game_height, c.leader.max_speed, 360, c.leader.max_rotation_speed,
                         c.game_width, c.game_height, c.follower.max_speed, 360, c.follower.max_rotation_speed),
                        dtype=np.float32)
        return low, high

    def ray_layout(self):
        """[(sensor name, offset into the rays vector, H, width)] in follower_sensors order."""
        out, off = [], 0
        for i, name in enumerate(self.ray_sensor_names):
            r = self.c.ray[i]
            w = 5 * r.lasers_count if r.compas else 4 * r.lasers_count if r.pad_sectors else r.lasers_count
            out.append((name, off, r.max_prev_obs, w))
            off += r.max_prev_obs * w
        return out

    @property
    def rays_per_env(self):
        return abi.rays_per_env(self.c)


# ---- presets baked into the reference's registered subclasses -------------------------------------
TEST_GAME_MANUAL_KWARGS = dict(  # ENV:1968-2011 (manual_control is what the batch simulator cannot do)
    add_obstacles=True, game_width=1500, game_height=1000, max_steps=15000, framerate=100, pixels_to_meter=50,
    obstacle_number=35, constant_follower_speed=False, min_distance=1, max_distance=4, max_dev=1, add_bear=True,
    bear_behind=False, multi_random_bears=False, move_bear_v4=True, bear_number=2, bear_max_speed=1.2,
    negative_speed=True, follower_max_speed=0.6, leader_max_speed=0.45, return_render_matrix=False,
    leader_speed_regime={0: [0.2, 1], 200: 1, 1000: [0.5, 1], 1500: 0.75, 2000: 0, 2500: 1, 3000: [0.5, 1],
                         4000: [0.0, 0.5], 5000: [0.4, 1]},
    leader_acceleration_regime={0: 0, 3100: 0.03, 4500: 0}, multiple_end_points=False, warm_start=0,
    frames_per_step=1, early_stopping={"max_distance_coef": 4, "low_reward": -300}, follower_sensors={})

TEST_GAME_MANUAL_GAZEBO_KWARGS = dict(  # ENV:2013-2107
    game_width=1500, game_height=1000, pixels_to_meter=10, step_grid=10, max_steps=30000, framerate=90,
    frames_per_step=5, min_distance=8, max_distance=15, max_dev=1, constant_follower_speed=False, warm_start=0,
    path_finding_iterations=15000, follower_size=(1, 1), leader_size=(4, 2), bear_size=(1.5, 1.5),
    follower_max_speed=2, leader_max_speed=1, negative_speed=True, bear_max_speed=1.2,
    follower_max_rotation_speed=28.65, leader_max_rotation_speed=28.65, follower_acceleration=1,
    leader_acceleration=1, leader_margin=1,
    leader_speed_regime={0: [0.2, 1], 200: 1, 1000: [0.5, 1], 1500: 0.75, 2300: 0, 2500: 1, 3000: [0.5, 1],
                         4000: [0.0, 0.5], 5000: [0.4, 1]},
    add_obstacles=True, obstacle_number=20, add_bear=True, bear_number=2, bear_behind=False,
    multi_random_bears=False, move_bear_v4=True, bridge_size=[140, 40], multiple_end_points=False,
    return_render_matrix=False, leader_acceleration_regime={0: 0, 3100: 0.03, 4500: 0},
    early_stopping={"max_distance_coef": 4, "low_reward": -300},
    follower_sensors={
        "LeaderPositionsTracker_v2": {
            "sensor_class": "LeaderPositionsTracker_v2", "eat_close_points": False, "generate_corridor": True,
            "saving_period": 8, "sensor_name": "LeaderPositionsTracker_v2", "start_corridor_behind_follower": True,
            "corridor_length": 250, "corridor_width": 30},
        "LeaderCorridor_lasers_all": {
            "sensor_name": "LeaderCorridor_lasers_all", "sensor_class": "LeaderCorridor_Prev_lasers_v2",
            "react_to_green_zone": True, "react_to_obstacles": True, "react_to_safe_corridor": True,
            "lasers_count": 12, "laser_length": 100, "max_prev_obs": 5, "use_prev_obs": True, "pad_sectors": False},
        "LeaderCorridor_lasers_obstacles": {
            "sensor_name": "LeaderCorridor_lasers_obstacles", "sensor_class": "LeaderCorridor_Prev_lasers_v2",
            "react_to_green_zone": False, "react_to_obstacles": True, "react_to_safe_corridor": False,
            "lasers_count": 24, "laser_length": 150, "max_prev_obs": 5, "use_prev_obs": True, "pad_sectors": False},
    })


def cfg3_sensors(corridor_rays=12, obstacle_rays=36, max_prev_obs=5):
    """The sensor stack BASELINE.json's config 3 names."""
    return {
        "LeaderPositionsTracker_v2": {
            "sensor_class": "LeaderPositionsTracker_v2", "sensor_name": "LeaderPositionsTracker_v2",
            "eat_close_points": False, "generate_corridor": True, "saving_period": 8,
            "start_corridor_behind_follower": True, "corridor_length": 350, "corridor_width": 75},
        "LeaderCorridor_Prev_lasers_v2": {
            "sensor_class": "LeaderCorridor_Prev_lasers_v2", "sensor_name": "LeaderCorridor_Prev_lasers_v2",
            "react_to_green_zone": True, "react_to_obstacles": True, "react_to_safe_corridor": True,
            "lasers_count": corridor_rays, "laser_length": 150, "max_prev_obs": max_prev_obs, "use_prev_obs": True,
            "pad_sectors": False},
        "LaserPrevSensor": {
            "sensor_class": "LeaderCorridor_Prev_lasers_v2", "sensor_name": "LaserPrevSensor",
            "react_to_green_zone": False, "react_to_obstacles": True, "react_to_safe_corridor": False,
            "lasers_count": obstacle_rays, "laser_length": 200, "max_prev_obs": max_prev_obs, "use_prev_obs": True,
            "pad_sectors": False, "first_laser_angle_offset": 0},
    }
