"""Observation post-processing of the shipped trained configurations, for ``gym_surface.Game`` and for the batch.

The reference stacks ``ContinuousObserveModifier_sensorPrev`` (utils/wrappers.py:169-221), ``SkipBadSeeds``
(utils/wrappers.py:814-825) and sometimes ``MyFrameStack`` (utils/wrappers.py:17-70) on its env.  The reference's OWN
wrapper classes run unmodified on ``gym_surface.Game`` (tests/test_reference_wrappers.py proves it on a golden
trace); the classes below are independent implementations of the same contracts -- same names, constructor
arguments, spaces and results -- for installations that do not carry the reference tree, built around one
table-driven normaliser that also serves the device tensors of ``FtlBatchEnv`` (``sensor_prev_observation``).
"""
import collections

import numpy as np

from .gym_surface import Box

# sensor classes whose (H, W) history blocks the sensorPrev contract normalises and concatenates (WRP:206, 214)
PREV_CLASSES = ("LeaderCorridor_Prev_lasers_v2", "LeaderCorridor_Prev_lasers_v3")
COMPAS_CLASS = "LeaderCorridor_lasers_compas"


def _sensor_prev_plan(follower_sensors):
    """[(dict key, width in features)] of the sensors the sensorPrev contract consumes, in dict order.

    Every entry must name its ``sensor_class`` (the reference indexes it unconditionally, WRP:181, so a config
    without it is a KeyError there as well); a sensor also counts when its dict KEY is one of the class names."""
    plan = []
    for key, conf in follower_sensors.items():
        cls = conf["sensor_class"]
        if cls in PREV_CLASSES:
            plan.append((key, conf["lasers_count"] * (4 if conf["pad_sectors"] else 1), True))
        elif cls == COMPAS_CLASS:
            plan.append((key, 5 * conf["lasers_count"], True))
        elif key in PREV_CLASSES or key == COMPAS_CLASS:
            plan.append((key, 0, False))   # consumed by observation() but never counted by the constructor upstream
    return plan


class _Delegate:
    """Minimal stand-in for gym.Wrapper: public attributes fall through to the wrapped env."""

    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def seed(self, seed=None):
        return self.env.seed(seed)

    def close(self):
        return self.env.close()

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


# names other code may import from here, as from gym
Wrapper = _Delegate


class ObservationWrapper(_Delegate):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info


class ContinuousObserveModifier_sensorPrev(ObservationWrapper):
    """obs dict -> float matrix [max_prev_obs, sum of sensor widths], each block ``clip(block / laser_length, 0, 1)``.

    ``action_values_range`` only re-declares the action space as [-1, 1]; like upstream (WRP:193-201, no ``step``
    override) this wrapper never rescales the actions it forwards."""

    def __init__(self, env, action_values_range=None, lz4_compress=False, max_prev_obs=0):
        super().__init__(env)
        self.max_prev_obs = max_prev_obs
        self.observations_list = None
        self._plan = _sensor_prev_plan(env.follower_sensors)
        self.features_number_num = sum(width for _, width, counted in self._plan if counted)
        bound = np.ones([max_prev_obs, self.features_number_num])
        self.observation_space = Box(-bound, bound)
        self.action_values_range = action_values_range
        if action_values_range is not None:
            lo, hi = action_values_range
            inner = env.action_space
            self.scale = (hi - lo) / (inner.high - inner.low)
            self.min = lo - inner.low * self.scale
            self.action_space = Box(low=-np.ones_like(inner.low), high=np.ones_like(inner.high), shape=inner.shape,
                                    dtype=inner.dtype)

    def observation(self, obs):
        blocks = []
        for key, _, _ in self._plan:
            block = obs[key]
            assert block.ndim == 2 and block.shape[0] == self.max_prev_obs
            blocks.append(np.clip(block / self.follower.sensors[key].laser_length, 0, 1))
        self.observations_list = np.concatenate(blocks, axis=1)
        return self.observations_list


class SkipBadSeeds(_Delegate):
    """``reset`` repeats until the route planner reached its target (``found_target_point``)."""

    def reset(self, **kwargs):
        while True:
            obs = self.env.reset(**kwargs)
            if self.env.found_target_point:
                return obs


class MyFrameStack(ObservationWrapper):
    """Rolling concatenation (axis 0) of the last ``framestack`` observations; ``reset`` fills the window with
    copies of the first one."""

    def __init__(self, env, framestack, lz4_compress=False):
        super().__init__(env)
        self.framestack = framestack
        self.lz4_compress = lz4_compress
        self.frames = collections.deque(maxlen=framestack)
        space = self.observation_space
        self.observation_space = Box(low=np.tile(space.low, framestack), high=np.tile(space.high, framestack),
                                     dtype=space.dtype)

    def observation(self, observation=None):
        assert len(self.frames) == self.framestack, (len(self.frames), self.framestack)
        return np.concatenate(self.frames)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.frames.append(obs)
        return self.observation(), reward, done, info

    def reset(self, **kwargs):
        first = self.env.reset(**kwargs)
        self.frames.extend([first] * self.framestack)
        return self.observation()


def sensor_prev_observation(batch_env):
    """The sensorPrev matrix for every env of an ``FtlBatchEnv``, on the device: [N, H, sum of widths].

    Only sensors the single-env wrapper would consume take part (history sensors of PREV_CLASSES; a flat
    ``LeaderCorridor_lasers(_v2)`` configured next to them is skipped, as WRP:206-213 skips it)."""
    import torch
    if batch_env.cfg.fused_sensor_prev:   # the ray kernel already wrote this matrix
        return batch_env.sensor_prev()
    gc = batch_env.gc
    wanted = {key for key, _, _ in _sensor_prev_plan(gc.follower_sensors)}
    feats = []
    for i, (name, off, h, w) in enumerate(gc.ray_layout()):
        if name not in wanted or gc.ray_sensor_flat[i]:
            continue
        # a tensor divisor: torch turns division by a python scalar into a multiplication by 1/L (1 ulp off numpy's)
        div = torch.tensor(gc.c.ray[i].laser_length, dtype=torch.float32, device=batch_env.rays.device)
        feats.append(torch.clamp(batch_env.rays[:, off:off + h * w].view(batch_env.n, h, w) / div, 0, 1))
    return torch.cat(feats, dim=2)
