"""The reference's gym surface, kept drop-in: registered env ids, ``Game`` and its preset subclasses,
``seed / reset / step`` with the old 4-tuple API, the action and observation spaces and the attributes the
wrappers read (follow_the_leader_continuous_env.py:44-105, 360-378, 429-543, 908-945, 1789-1824, 1963-2169).

A ``Game`` here is a single-env view over a batch of one: ``reset()`` draws the scenario on the host exactly
like the reference does (scenario_gen.py), uploads it, and ``step()`` goes through ``ftl_step_host``.  For
throughput use ``FtlBatchEnv`` (batch_env.py) -- same kernels, thousands of envs per launch.

If the real ``gym`` package is importable its ``Env``/``spaces``/registry are used; otherwise the small
stand-ins below provide the same calls (``make``, ``register``, ``spaces.Box``, ``spaces.Discrete``).
"""
import random

import numpy as np

from . import abi, capi, scenario_gen
from .config import GameConfig, TEST_GAME_MANUAL_GAZEBO_KWARGS, TEST_GAME_MANUAL_KWARGS
from .scenario import ScenarioPool

try:  # pragma: no cover - gym is not in the build image
    import gym as _gym
    from gym.spaces import Box, Discrete
    Env = _gym.Env
    _HAVE_GYM = True
except Exception:  # noqa: BLE001
    _HAVE_GYM = False

    class Env:
        metadata = {}
        action_space = None
        observation_space = None

        @property
        def unwrapped(self):
            return self

        def close(self):
            return None

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            self.shape = tuple(shape) if shape is not None else low.shape
            self.low = np.broadcast_to(low, self.shape).copy()
            self.high = np.broadcast_to(high, self.shape).copy()
            self.dtype = np.dtype(dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, self.dtype)

    class Discrete:
        def __init__(self, n):
            self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            return int(self._rng.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return "Discrete(%d)" % self.n


class _SensorView:
    """What wrappers read from ``env.follower.sensors[name]`` (WRP:212)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class _RobotView:
    """``env.leader`` / ``env.follower``: position, direction, speed ... read lazily from device state."""

    def __init__(self, game, which):
        self._game, self._which = game, which
        self.sensors = {}

    def _state(self):
        return self._game._state_record()[self._which]

    position = property(lambda s: np.array(s._state()["pos"], dtype=np.float32))
    direction = property(lambda s: float(s._state()["dir"]))
    speed = property(lambda s: float(s._state()["speed"]))
    rotation_speed = property(lambda s: float(s._state()["rot_speed"]))
    rectangle = property(lambda s: tuple(int(v) for v in s._state()["rect"]))

    @property
    def max_speed(self):
        rc = self._game.gc.c.leader if self._which == "leader" else self._game.gc.c.follower
        return rc.max_speed


class Game(Env):
    """Drop-in for the reference ``Game`` (same constructor keywords, ENV:45-105)."""

    metadata = {"render.modes": ["rgb_array"]}

    def __init__(self, device=0, lib=None, **kwargs):
        self.gc = GameConfig(route_cap=kwargs.pop("route_cap", None), **kwargs)
        g, c = self.gc.kwargs, self.gc.c
        self._lib, self._device = lib, device
        self._env = None
        self._caps = None
        self._trajectory_arg = g["trajectory"]
        # public attributes of the reference object
        self.follower_sensors = g["follower_sensors"]
        self.PIXELS_TO_METER = g["pixels_to_meter"]
        self.DISPLAY_WIDTH, self.DISPLAY_HEIGHT = c.game_width, c.game_height
        self.min_distance, self.max_distance, self.max_dev = c.min_distance, c.max_distance, c.max_dev
        self.frames_per_step, self.max_steps, self.warm_start = c.frames_per_step, c.max_steps, c.warm_start
        self.leader_pos_epsilon = g["leader_pos_epsilon"]
        self.early_stopping = g["early_stopping"]
        self.random_frames_per_step = self.gc.random_frames_per_step
        if self.random_frames_per_step is not None:   # ENV:405: drawn once here, redrawn after every step (ENV:939-940)
            self.frames_per_step = int(np.random.randint(*self.random_frames_per_step))
        self.found_target_point = False
        self.trajectory = g["trajectory"]
        self.finish_point = (10, 10)
        self.multiple_end_points = bool(g["multiple_end_points"])
        if self.multiple_end_points:
            self.finish_point2 = (1490, 990)          # ENV:246
        self.simulation_number = 0
        self.done = False
        if self.gc.discrete_action_space:
            self.action_space = Discrete(5)
        else:
            lo, hi = self.gc.action_bounds()
            self.action_space = Box(lo, hi, shape=lo.shape, dtype=np.float32) if self.gc.constant_follower_speed \
                else Box(lo, hi)
        lo, hi = self.gc.observation_bounds()
        self.observation_space = Box(low=lo, high=hi)
        self.leader, self.follower = _RobotView(self, "leader"), _RobotView(self, "follower")
        for i, name in enumerate(self.gc.ray_sensor_names):
            r = c.ray[i]
            self.follower.sensors[name] = _SensorView(laser_length=r.laser_length, lasers_count=r.lasers_count,
                                                      max_prev_obs=r.max_prev_obs, pad_sectors=bool(r.pad_sectors))
        if c.tracker_enabled:
            self.follower.sensors["LeaderPositionsTracker_v2"] = _SensorView(
                saving_period=c.saving_period, corridor_length=c.corridor_length, corridor_width=c.corridor_width,
                generate_corridor=True)
        self._cached_state = None

    # ---- gym API ------------------------------------------------------------------------------------------
    def seed(self, seed_value):  # ENV:429-432
        random.seed(seed_value)
        np.random.seed(seed_value)
        return

    def reset(self, scenario=None):
        """Game.reset (ENV:434-543).  ``scenario`` (an extension) injects a ready scenario_gen.Scenario instead of
        drawing one, e.g. to replay a layout exported from the reference."""
        sc = scenario if scenario is not None else scenario_gen.generate(self.gc, trajectory=self._trajectory_arg)
        self.trajectory = list(sc.route)
        self.finish_point = sc.finish_point
        if self.multiple_end_points and sc.finish_points is not None:      # ENV:472-482
            self.finish_point, self.finish_point2, self.finish_point3 = sc.finish_points
        self.found_target_point = bool(sc.found_target_point)
        self._ensure_env(len(sc.route), len(sc.static_rects))
        pool = ScenarioPool(1, self.gc.c.static_cap, self.gc.c.route_cap)
        pool.set(0, sc.static_rects, sc.route, sc.leader_pos, sc.leader_dir, sc.follower_pos, sc.follower_dir,
                 sc.found_target_point)
        self._env.upload_scenarios(pool)
        out = self._env.reset(scenario_ids=np.zeros(1, np.int32))
        self.simulation_number += 1
        self.done = False
        self._cached_state = None
        return self._obs(out)

    def step(self, action):
        if self._env is None:
            raise RuntimeError("reset() must be called before step()")
        if self.gc.discrete_action_space:
            a = np.asarray(action)
            if a.ndim == 2:
                assert a.shape[0] == 1 and a.shape[1] == 1   # ENV:919-921
                a = a[0, 0]
            a = np.array([int(a)], np.int32)
        else:
            a = np.asarray(action, dtype=np.float32).reshape(1, -1)
        if self.random_frames_per_step is not None:
            out = self._env.step(a, frames=np.array([self.frames_per_step], np.int32))
            self.frames_per_step = int(np.random.randint(*self.random_frames_per_step))
        else:
            out = self._env.step(a)
        self._cached_state = None
        st = out.status[0]
        info = {"mission_status": abi.MISSION_STATUS[int(st[0])], "agent_status": abi.AGENT_STATUS[int(st[1])],
                "leader_status": abi.LEADER_STATUS[int(st[2])]}
        self.done = bool(out.done[0])
        return self._obs(out), float(out.reward[0]), self.done, info

    def render(self, mode="rgb_array", scale=1, **kwargs):
        """ENV:1196-1202 with ``return_render_matrix=True``: the frame as uint8 [game_height, game_width, 3] (the layout
        of ``np.transpose(surfarray.array3d(display), (1, 0, 2))``), rasterised on the device by ``ftl_render`` -- the
        layers and colours of ``_show_tick`` (ENV:1229-1302) with hit boxes instead of sprites and without the text
        lines.  There is no window: ``mode="human"`` raises."""
        if mode != "rgb_array":
            raise NotImplementedError("only mode='rgb_array' is rendered (no pygame window in the accelerated path)")
        if self._env is None:
            raise RuntimeError("render() before reset()")
        return self._env.render(0, 1, scale=scale)[0]

    def close(self):
        if self._env is not None:
            self._env.close()
            self._env = None

    # ---- attributes other code reads (RUN:93, WRP:823) ------------------------------------------------------
    @property
    def overall_reward(self):
        return float(self._state_record()["overall_reward"])

    @property
    def step_count(self):
        return int(self._state_record()["step_count"])

    @property
    def leader_factual_trajectory(self):
        st = self._env.get_state()
        n = int(st.env[0]["trail_len"])
        return [np.array(p, np.float32) for p in st.trail[0, :n]]

    # ---- internals ------------------------------------------------------------------------------------------------
    def _ensure_env(self, route_len, n_static):
        c = self.gc.c
        need_route = max(int(route_len), 2)
        need_static = max(int(n_static), 1)
        if self._env is None or need_route > c.route_cap or need_static > c.static_cap:
            if self._env is not None:
                self._env.close()
            if need_route > c.route_cap:
                c.route_cap = max(128, 1 << (need_route - 1).bit_length())
            if need_static > c.static_cap:
                c.static_cap = need_static
            self._env = capi.HostEnv(self.gc, 1, device=self._device, lib=self._lib)

    def _state_record(self):
        if self._cached_state is None:
            self._cached_state = self._env.get_state().env[0]
        return self._cached_state

    def _obs(self, out):  # ENV:1789-1810
        obs = {"numerical_features": out.numerical_features[0].copy(),
               "leader_target_point": (int(out.leader_target[0, 0]), int(out.leader_target[0, 1]))}
        if self.gc.c.tracker_enabled:
            obs["LeaderPositionsTracker_v2"] = self._tracker_obs()   # CLS:263-286 puts the tracker's tuple in the dict
        for k, (name, off, h, w) in enumerate(self.gc.ray_layout()):
            block = out.rays[0, off:off + h * w]
            # the sensors without history return (R,) (SEN:724-726, 805-807), the history sensors (H, R)
            obs[name] = block.copy() if self.gc.ray_sensor_flat[k] else block.reshape(h, w).copy()
        if self.gc.follower_info_name is not None:   # SEN:834-842
            obs[self.gc.follower_info_name] = out.follower_info[0].copy()
        if self.gc.track_vector_name is not None:    # SEN:365-380
            obs[self.gc.track_vector_name] = out.track_vectors[0].copy()
        if self.gc.radar_name is not None:           # SEN:425-461
            obs[self.gc.radar_name] = out.radar[0].copy()
        if self.gc.laser_name is not None:           # SEN:63-136
            obs[self.gc.laser_name] = out.laser[0].copy()
        return obs

    def _tracker_obs(self):
        st = self._env.get_state()
        e, cap = st.env[0], self.gc.c.corridor_cap
        idx = [k % cap for k in range(int(e["ring_tail"]), int(e["ring_head"]))]
        hist = [st.hist[0, k].copy() for k in idx]
        corridor = [[st.corridor[0, k, :2].astype(np.float64), st.corridor[0, k, 2:].astype(np.float64)] for k in idx]
        return hist, corridor


# ---- preset subclasses, ENV:1963-2129 --------------------------------------------------------------------------
class TestGameAuto(Game):
    __test__ = False

    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class TestGameManual(Game):
    """The reference preset hard-codes manual_control=True (keyboard teleop); constructing it raises here."""
    __test__ = False

    def __init__(self, **kwargs):
        super().__init__(manual_control=True, **dict(TEST_GAME_MANUAL_KWARGS, **kwargs))


class TestGameManual_gazebo(Game):
    __test__ = False

    def __init__(self, **kwargs):
        super().__init__(manual_control=True, **dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, **kwargs))


class TestGameBaseAlgoNoObst(Game):
    __test__ = False

    def __init__(self):
        super().__init__(manual_control=False, add_obstacles=False, game_width=1500, game_height=1000,
                         early_stopping={"max_distance_coef": 1.2, "low_reward": -100})


class TestGameBaseAlgoObst(Game):
    __test__ = False

    def __init__(self):
        super().__init__(manual_control=False, add_obstacles=True, game_width=1500, game_height=1000,
                         early_stopping={"max_distance_coef": 1.2, "low_reward": -100},
                         follower_sensors={"GreenBoxBorderSensor": {"sensor_range": 2, "available_angle": 180,
                                                                    "angle_step": 45}})


class TestGameNEAT(Game):
    __test__ = False

    def __init__(self):
        super().__init__(manual_control=False, add_obstacles=False,
                         early_stopping={"max_distance_coef": 1.2, "low_reward": -100}, discrete_action_space=True)


# ---- registry, ENV:2132-2169 -----------------------------------------------------------------------------------
_MOD = "continiousenvironment_follower_leader_b200.gym_surface"
REGISTERED = {
    "Test-Cont-Env-Auto-v0": _MOD + ":TestGameAuto",
    "Test-Cont-Env-Manual-v0": _MOD + ":TestGameManual",
    "Test-Cont-Env-Manual-gazebo-v0": _MOD + ":TestGameManual_gazebo",
    # the reference registers these two against classes that do not exist (ENV:2149-2156); kept, and as there,
    # making them fails
    "Test-Cont-Env-Manual-hardcore-v0": _MOD + ":TestGameManual_hardcore",
    "Test-Cont-Env-Manual-gazebo-hardcore-v0": _MOD + ":TestGameManual_gazebo_hardcore",
    "Test-Cont-Env-Auto-Follow-no-obstacles-v0": _MOD + ":TestGameBaseAlgoNoObst",
    "Test-Cont-Env-Auto-Follow-with-obstacles-v0": _MOD + ":TestGameBaseAlgoObst",
    "Test-Game-Neat-v0": _MOD + ":TestGameNEAT",
}


def register_all():
    if _HAVE_GYM:  # pragma: no cover
        from gym.envs.registration import register
        for env_id, entry in REGISTERED.items():
            try:
                register(id=env_id, entry_point=entry, reward_threshold=10000)
            except Exception:  # already registered
                pass


def make(env_id, **kwargs):
    """gym.make for the registered ids (works with or without the gym package)."""
    if env_id not in REGISTERED:
        raise KeyError("No registered env with id: %s" % env_id)
    import importlib
    mod_name, attr = REGISTERED[env_id].split(":")
    cls = getattr(importlib.import_module(mod_name), attr)   # AttributeError for the two dangling ids, as upstream
    return cls(**kwargs)


register_all()
