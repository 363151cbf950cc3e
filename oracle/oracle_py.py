"""ctypes front-end of the CPU oracle (oracle/ftl_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from continiousenvironment_follower_leader_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libftl_oracle.so")
    src = os.path.join(_HERE, "ftl_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ftl.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libftl_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ftl_oracle_create.restype = C.c_void_p
        L.ftl_oracle_create.argtypes = [C.POINTER(abi.FtlConfig), C.c_int, C.c_int64, C.c_int]
        L.ftl_oracle_destroy.argtypes = [C.c_void_p]
        L.ftl_oracle_upload_scenarios.argtypes = [C.c_void_p, C.POINTER(abi.FtlScenarioPool)]
        L.ftl_oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(abi.FtlOutputs)]
        L.ftl_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(abi.FtlOutputs)]
        L.ftl_oracle_step_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(abi.FtlOutputs)]
        L.ftl_oracle_get_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.FtlStateBuffers)]
        L.ftl_oracle_set_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.FtlStateBuffers)]
        L.ftl_oracle_rays_per_env.argtypes = [C.c_void_p]
        assert L.ftl_oracle_sizeof_env_state() == C.sizeof(abi.FtlEnvState)
        assert L.ftl_oracle_sizeof_config() == C.sizeof(abi.FtlConfig)
        _LIB = L
    return _LIB


from continiousenvironment_follower_leader_b200.capi import HostOutputs, HostState  # noqa: E402,F401


class OracleEnv:
    def __init__(self, game_config, n_envs, env_id_base=0, n_threads=1):
        self.gc = game_config
        self.cfg = game_config.c
        self.n = n_envs
        self._L = lib()
        self._h = self._L.ftl_oracle_create(C.byref(self.cfg), n_envs, env_id_base, n_threads)
        if not self._h:
            raise ValueError("oracle rejected the configuration")
        self.out = HostOutputs(n_envs, abi.rays_per_env(self.cfg),
                               follower_info=getattr(game_config, "follower_info_name", None) is not None,
                               track_vector_len=self.cfg.track_vector_len, radar_sectors=self.cfg.radar_sectors,
                               laser_shape=getattr(game_config, "laser_shape", None))

    def close(self):
        if self._h:
            self._L.ftl_oracle_destroy(self._h)
            self._h = None

    __del__ = close

    def upload_scenarios(self, pool):
        rc = self._L.ftl_oracle_upload_scenarios(self._h, C.byref(pool.c_struct()))
        if rc:
            raise ValueError("upload_scenarios failed: %d" % rc)
        self._pool = pool

    def reset(self, mask=None, scenario_ids=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        s = None if scenario_ids is None else np.ascontiguousarray(scenario_ids, np.int32)
        rc = self._L.ftl_oracle_reset(self._h, abi.ptr(m), abi.ptr(s), C.byref(self.out.c))
        if rc:
            raise RuntimeError("oracle reset failed: %d" % rc)
        return self.out

    def step(self, actions, frames=None, regime_draws=None):
        """frames: int32[N] frames of this step per env; regime_draws: float64[N, frames_per_step] (FtlStepInputs)."""
        if self.cfg.action_mode == abi.ACTION_DISCRETE:
            a = np.ascontiguousarray(actions, np.int32)
        else:
            a = np.ascontiguousarray(actions, np.float32)
        if frames is not None or regime_draws is not None:
            f = None if frames is None else np.ascontiguousarray(frames, np.int32)
            d = None if regime_draws is None else np.ascontiguousarray(regime_draws, np.float64)
            assert d is None or d.shape == (self.n, self.cfg.frames_per_step)
            rc = self._L.ftl_oracle_step_ex(self._h, abi.ptr(a), abi.ptr(f), abi.ptr(d), C.byref(self.out.c))
        else:
            rc = self._L.ftl_oracle_step(self._h, abi.ptr(a), C.byref(self.out.c))
        if rc:
            raise RuntimeError("oracle step failed: %d" % rc)
        return self.out

    def get_state(self, first=0, n=None):
        n = self.n - first if n is None else n
        st = HostState(n, self.cfg)
        self._L.ftl_oracle_get_state(self._h, first, n, C.byref(st.c))
        return st

    def set_state(self, st, first=0):
        self._L.ftl_oracle_set_state(self._h, first, len(st.env), C.byref(st.c))
