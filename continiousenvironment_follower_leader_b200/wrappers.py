"""The wrappers the shipped trained configurations stack on the env (utils/wrappers.py):
``ContinuousObserveModifier_sensorPrev`` (WRP:169-221) and ``SkipBadSeeds`` (WRP:814-825), plus
``MyFrameStack`` (WRP:17-70).  They work on ``gym_surface.Game`` exactly as upstream; for batched
training ``sensor_prev_observation`` applies the same normalise-and-concatenate to the device tensors.
"""
from collections import deque

import numpy as np

from .gym_surface import Box

PREV_CLASSES = ("LeaderCorridor_Prev_lasers_v2", "LeaderCorridor_Prev_lasers_v3")


class Wrapper:
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        return self.observation(observation), reward, done, info


class ContinuousObserveModifier_sensorPrev(ObservationWrapper):
    """History-sensor features, each divided by its laser_length, clipped to [0, 1], concatenated on axis 1."""

    def __init__(self, env, action_values_range=None, lz4_compress=False, max_prev_obs=0):
        super().__init__(env)
        self.observations_list = None
        features_number = 0
        self.max_prev_obs = max_prev_obs
        for sensor_name, sensor_config in env.follower_sensors.items():
            if sensor_config["sensor_class"] in PREV_CLASSES:   # KeyError without "sensor_class", as upstream (WRP:181)
                if sensor_config["pad_sectors"]:
                    features_number += 4 * sensor_config["lasers_count"]
                else:
                    features_number += sensor_config["lasers_count"]
            if sensor_config["sensor_class"] == "LeaderCorridor_lasers_compas":
                features_number += 5 * sensor_config["lasers_count"]
        self.features_number_num = features_number
        self.observation_space = Box(-np.ones([self.max_prev_obs, features_number]),
                                     np.ones([self.max_prev_obs, features_number]))
        self.action_values_range = action_values_range
        if self.action_values_range is not None:   # declared only: this wrapper never rescales in step (WRP:193-201)
            low_bound, high_bound = self.action_values_range
            self.scale = (high_bound - low_bound) / (env.action_space.high - env.action_space.low)
            self.min = low_bound - env.action_space.low * self.scale
            self.action_space = Box(low=-np.ones_like(env.action_space.low), high=np.ones_like(env.action_space.high),
                                    shape=env.action_space.shape, dtype=env.action_space.dtype)

    def observation(self, obs):
        features_list = []
        for sensor_name in self.follower_sensors.keys():
            sensor_config = self.follower_sensors[sensor_name]
            if sensor_name in PREV_CLASSES or sensor_config["sensor_class"] in PREV_CLASSES:
                corridor_obs = obs[sensor_name]
                assert len(corridor_obs.shape) == 2
                assert corridor_obs.shape[0] == self.max_prev_obs
                corridor_obs = np.clip(corridor_obs / self.follower.sensors[sensor_name].laser_length, 0, 1)
                features_list.append(corridor_obs)
        self.observations_list = np.concatenate(features_list, axis=1)
        return self.observations_list


class SkipBadSeeds(Wrapper):
    """Re-reset until the planner reached the target (WRP:814-825)."""

    def reset(self, **kwargs):
        observation = self.env.reset(**kwargs)
        while not self.env.found_target_point:
            observation = self.env.reset(**kwargs)
        return observation


class MyFrameStack(ObservationWrapper):
    def __init__(self, env, framestack, lz4_compress=False):
        super().__init__(env)
        self.framestack = framestack
        self.frames = deque(maxlen=framestack)
        low = np.tile(self.observation_space.low[...], framestack)
        high = np.tile(self.observation_space.high[...], framestack)
        self.observation_space = Box(low=low, high=high, dtype=self.observation_space.dtype)

    def observation(self, observation=None):
        assert len(self.frames) == self.framestack, (len(self.frames), self.framestack)
        return np.concatenate(self.frames)

    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        self.frames.append(observation)
        return self.observation(), reward, done, info

    def reset(self, **kwargs):
        observation = self.env.reset(**kwargs)
        [self.frames.append(observation) for _ in range(self.framestack)]
        return self.observation()


def sensor_prev_observation(batch_env):
    """Batched ContinuousObserveModifier_sensorPrev.observation on device tensors: [N, H, sum of widths]."""
    import torch
    if batch_env.cfg.fused_sensor_prev:   # the ray kernel already wrote this matrix
        return batch_env.sensor_prev()
    feats = []
    layout = batch_env.gc.ray_layout()
    for i, (name, off, h, w) in enumerate(layout):
        L = batch_env.gc.c.ray[i].laser_length
        # a tensor divisor: torch turns division by a python scalar into a multiplication by 1/L (1 ulp off numpy's)
        div = torch.tensor(L, dtype=torch.float32, device=batch_env.rays.device)
        feats.append(torch.clamp(batch_env.rays[:, off:off + h * w].view(batch_env.n, h, w) / div, 0, 1))
    return torch.cat(feats, dim=2)
