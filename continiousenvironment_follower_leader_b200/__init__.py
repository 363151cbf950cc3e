"""B200-native batched simulator for the 2D "follow the leader" environment.

Drop-in for the hot path of sag111/ContiniousEnvironment_Follower_Leader
(src/continuous_grid_arctic: Game.step and what it calls).  The compute lives in hand-written
sm_100a CUDA kernels behind the C-ABI in include/ftl.h (libftl.so); this package is the host side.
There is no CPU fallback: creating an environment without the compiled library raises.
"""
from .config import GameConfig, cfg3_sensors  # noqa: F401

__all__ = ["GameConfig", "cfg3_sensors"]
