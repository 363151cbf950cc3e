"""Run the UNMODIFIED reference simulator head-less -- TEST INFRASTRUCTURE ONLY.

This module exists so that golden traces can be exported from the real reference code
(oracle/gen_golden.py) and so that the C restatement (oracle/ftl_oracle.c) can be pinned against
it.  It only works where /root/reference exists (the build container); nothing that runs on the GPU
box may import it, and nothing in the product package does.

What it does:
  1. puts oracle/shims (gym + pygame stand-ins) and /root/reference on sys.path and imports
     ``src.continuous_grid_arctic.follow_the_leader_continuous_env`` unchanged;
  2. wraps ``random.randrange`` so integral floats are accepted, as on the reference's Python 3.7
     (the reference passes floats at follow_the_leader_continuous_env.py:549; Python >= 3.12 rejects
     them);
  3. offers helpers that feed actions as *Python floats* (numpy >= 2 would otherwise run the
     controllers in float32) and that dump the complete simulator state.
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("FTL_REFERENCE_ROOT", "/root/reference")

_env_mod = None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "continuous_grid_arctic"))


def load_reference():
    """Import the reference env module under the shims; returns the module."""
    global _env_mod
    if _env_mod is not None:
        return _env_mod
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    shims = os.path.join(_HERE, "shims")
    for p in (REFERENCE_ROOT, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    _patch_randrange()
    import importlib
    _env_mod = importlib.import_module("src.continuous_grid_arctic.follow_the_leader_continuous_env")
    return _env_mod


def _patch_randrange():
    if getattr(random, "_ftl_patched", False):
        return
    inst = random._inst
    orig = inst.randrange

    def _as_int(v):
        if isinstance(v, (float, np.floating)):
            iv = int(v)
            if iv != v:
                raise ValueError("non-integer arg for randrange()")
            return iv
        return v

    def randrange(start, stop=None, step=1):
        return orig(_as_int(start), None if stop is None else _as_int(stop), _as_int(step))

    random.randrange = randrange
    random._ftl_patched = True


@contextlib.contextmanager
def quiet():
    """The reference prints at every reset/crash (ENV:444, 1069, 1902)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def make_env(env_id="Test-Cont-Env-Auto-v0", **kwargs):
    load_reference()
    import gym
    with quiet():
        return gym.make(env_id, **kwargs)


class StepInputRecorder:
    """Records, per env.step, what the reference drew from its global RNGs inside the step: the frames_per_step the step
    ran with (random_frames_per_step, ENV:939-940) and, per frame, the random() behind random.uniform of a list-valued
    leader_speed_regime entry (ENV:1155-1156) -- the FtlStepInputs a replay needs (include/ftl.h)."""

    def __init__(self, env, frames_cap):
        self.env, self.cap = env, int(frames_cap)
        self.frames, self.draws = [], []
        self._frame = 0
        self._cur = None
        inner_frame_step = env.frame_step

        def frame_step(action):
            out = inner_frame_step(action)
            self._frame += 1
            return out

        env.frame_step = frame_step
        rec = self

        def uniform(a, b):   # CPython: a + (b - a) * self.random()
            u = random.random()
            if rec._cur is not None and rec._frame < rec.cap:
                rec._cur[rec._frame] = u
            return a + (b - a) * u

        random.uniform = uniform

    def begin_step(self):
        self._frame = 0
        self._cur = np.zeros(self.cap, np.float64)
        self.frames.append(int(self.env.frames_per_step))

    def end_step(self):
        self.draws.append(self._cur)
        self._cur = None


def py_action(a):
    """Action as python floats holding exactly the float32 values the GPU path receives."""
    return [float(np.float32(x)) for x in a]


def step(env, action):
    with quiet():
        return env.step(py_action(action))


def reset(env, seed=None):
    with quiet():
        if seed is not None:
            env.seed(seed)
        return env.reset()


# ------------------------------------------------------------------------------------------------
# state export
# ------------------------------------------------------------------------------------------------
def robot_state(r):
    """f64[8] scalars + i32[4] rect of one reference robot object (CLS:59-107)."""
    rect = r.rectangle
    return (np.array([r.position[0], r.position[1], r.direction, r.speed, r.rotation_speed,
                      r.rotation_direction, r.desirable_speed, r.desirable_rotation_speed,
                      r.desirable_rotation_direction], dtype=np.float64),
            np.array([rect.x, rect.y, rect.w, rect.h], dtype=np.int32))


def static_rects(env):
    """i32[S,4] of walls + rocks in game_object_list order (ENV:675-677)."""
    out = []
    for o in env.game_object_list:
        if o is env.leader or o is env.follower:
            continue
        r = o.rectangle
        out.append((r.x, r.y, r.w, r.h))
    return np.array(out, dtype=np.int32).reshape(-1, 4)
