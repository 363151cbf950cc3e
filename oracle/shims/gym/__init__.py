"""Minimal stand-in for the gym 0.21-era API -- TEST INFRASTRUCTURE ONLY.

Lets the unmodified reference (which does ``import gym`` and registers its env ids at import time,
follow_the_leader_continuous_env.py:12-14, 2132-2169) import in a container without gym.  The
product package ships its own registry (continiousenvironment_follower_leader_b200/gym_surface.py)
and never imports this.
"""
from . import spaces  # noqa: F401
from .core import Env, Wrapper, ObservationWrapper, ActionWrapper, RewardWrapper  # noqa: F401
from .envs.registration import make, register  # noqa: F401
from . import envs  # noqa: F401
