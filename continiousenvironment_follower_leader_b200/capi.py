"""ctypes binding of libftl.so (include/ftl.h) and the host-buffer environment built on it.

``HostEnv`` is the call a gym-style user makes with numpy data: actions come from host memory, the
observation/reward/done arrays land in host memory (``ftl_step_host``: H2D copy, the fused step
kernels, D2H copies, one stream synchronise).  The device-resident path used for throughput lives in
``batch_env.py``.

There is no CPU fallback: if libftl.so has not been built (``python -m
continiousenvironment_follower_leader_b200.build``), loading raises ``FtlLibraryMissing``.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_PKG = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_PKG, "csrc", "libftl.so")


class FtlLibraryMissing(ImportError):
    pass


class FtlError(RuntimeError):
    pass


_LIBS = {}


def load(path=None):
    """dlopen a library exporting the include/ftl.h entry points and declare their signatures."""
    path = os.path.abspath(path or os.environ.get("FTL_LIB") or DEFAULT_LIB)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise FtlLibraryMissing(
            "%s not found: build the CUDA extension first (python -m continiousenvironment_follower_leader_b200.build); "
            "there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.ftl_last_error.restype = C.c_char_p
    L.ftl_create.argtypes = [C.POINTER(abi.FtlConfig), i32, i32, i64, C.POINTER(vp)]
    L.ftl_destroy.argtypes = [vp]
    L.ftl_rays_per_env.argtypes = [vp]
    L.ftl_num_envs.argtypes = [vp]
    L.ftl_upload_scenarios.argtypes = [vp, C.POINTER(abi.FtlScenarioPool)]
    L.ftl_reset_host.argtypes = [vp, vp, vp, C.POINTER(abi.FtlOutputs), vp]
    L.ftl_step_host.argtypes = [vp, vp, C.POINTER(abi.FtlOutputs), vp]
    for name, argtypes, restype in (("ftl_step_host_begin", [vp, vp, C.POINTER(abi.FtlOutputs), vp], C.c_int),
                                    ("ftl_step_host_wait", [vp], C.c_int),
                                    ("ftl_host_stream", [vp], vp)):
        if hasattr(L, name):   # not in the host-compiled test harness
            getattr(L, name).argtypes = argtypes
            getattr(L, name).restype = restype
    L.ftl_get_state.argtypes = [vp, i32, i32, C.POINTER(abi.FtlStateBuffers)]
    L.ftl_set_state.argtypes = [vp, i32, i32, C.POINTER(abi.FtlStateBuffers)]
    L.ftl_generate_scenarios.argtypes = [C.POINTER(abi.FtlScenarioGenConfig), vp, i32, C.POINTER(abi.FtlScenarioPool), i32]
    for name, argtypes, restype in (
            ("ftl_reset", [vp, vp, vp, C.POINTER(abi.FtlOutputs), vp], C.c_int),
            ("ftl_step", [vp, vp, C.POINTER(abi.FtlOutputs), vp], C.c_int),
            ("ftl_stats", [vp, vp, i32, vp], C.c_int),
            ("ftl_launch_count", [vp], i64),
            ("ftl_profile", [vp, i32], C.c_int),
            ("ftl_profile_read", [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)], C.c_int)):
        if hasattr(L, name):  # the host-compiled test harness only has the *_host entry points
            getattr(L, name).argtypes = argtypes
            getattr(L, name).restype = restype
    _LIBS[path] = L
    return L


def check(L, rc, what):
    if rc != 0:
        msg = L.ftl_last_error()
        msg = msg.decode() if msg else ""
        if rc == abi.FTL_ERR_INVALID:
            raise ValueError("%s: %s" % (what, msg))
        raise FtlError("%s failed (%d): %s" % (what, rc, msg))


class HostOutputs:
    """numpy buffers laid out as FtlOutputs."""

    def __init__(self, n, rays_per_env, alloc=None, follower_info=False, track_vector_len=0, radar_sectors=0):
        alloc = alloc or (lambda shape, dtype: np.zeros(shape, dtype))
        self.numerical_features = alloc((n, 10), np.float32)
        self.leader_target = alloc((n, 2), np.int32)
        self.rays = alloc((n, max(rays_per_env, 1)), np.float32)
        self.reward = alloc((n,), np.float32)
        self.done = alloc((n,), np.uint8)
        self.status = alloc((n, 4), np.uint8)
        # the optional sensors' outputs exist only when the configuration has them (NULL pointers otherwise)
        self.follower_info = alloc((n, 2), np.float32) if follower_info else None
        self.track_vectors = alloc((n, track_vector_len, 2), np.float32) if track_vector_len else None
        self.radar = alloc((n, radar_sectors), np.float32) if radar_sectors else None
        self.c = abi.FtlOutputs(abi.ptr(self.numerical_features), abi.ptr(self.leader_target), abi.ptr(self.rays),
                                abi.ptr(self.reward), abi.ptr(self.done), abi.ptr(self.status),
                                None if self.follower_info is None else abi.ptr(self.follower_info),
                                None if self.track_vectors is None else abi.ptr(self.track_vectors),
                                None if self.radar is None else abi.ptr(self.radar))

    def nbytes(self):
        return sum(a.nbytes for a in (self.numerical_features, self.leader_target, self.rays, self.reward, self.done,
                                      self.status, self.follower_info, self.track_vectors, self.radar) if a is not None)


class HostState:
    """numpy buffers laid out as FtlStateBuffers."""

    def __init__(self, n, cfg):
        self.env = np.zeros(n, abi.ENV_STATE_DTYPE)
        self.trail = np.zeros((n, cfg.trail_cap, 2), np.float32)
        self.hist = np.zeros((n, cfg.corridor_cap, 2), np.float64)
        self.corridor = np.zeros((n, cfg.corridor_cap, 4), np.float32)
        self.c = abi.FtlStateBuffers(abi.ptr(self.env), abi.ptr(self.trail), abi.ptr(self.hist), abi.ptr(self.corridor))


def pinned_alloc(shape, dtype):
    """Page-locked host array (through torch) so the D2H/H2D copies of the host path run at full PCIe rate."""
    import torch
    t = torch.empty(tuple(int(x) for x in shape), dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
    t.zero_()
    a = t.numpy()
    return a, t


class HostEnv:
    """N environments behind host (numpy) buffers; mirrors Game.reset/Game.step for a batch."""

    def __init__(self, game_config, n_envs, device=0, env_id_base=0, lib=None, pinned=False, own_stream=False):
        """own_stream: run on a non-blocking stream owned by the handle (ftl_host_stream) instead of the legacy
        default stream, so that several HostEnv objects -- halves of a batch, vector-env workers in threads -- overlap:
        one's kernels run while another's results cross PCIe."""
        self.gc = game_config
        self.cfg = game_config.c
        self.n = int(n_envs)
        self._L = lib if lib is not None else load()
        self._h = C.c_void_p()
        check(self._L, self._L.ftl_create(C.byref(self.cfg), self.n, int(device), int(env_id_base), C.byref(self._h)),
              "ftl_create")
        self._keep = []
        alloc = None
        if pinned:
            def alloc(shape, dtype):
                a, t = pinned_alloc(shape, dtype)
                self._keep.append(t)
                return a
        self.out = HostOutputs(self.n, abi.rays_per_env(self.cfg), alloc,
                               follower_info=getattr(game_config, "follower_info_name", None) is not None,
                               track_vector_len=self.cfg.track_vector_len, radar_sectors=self.cfg.radar_sectors)
        if self.cfg.action_mode == abi.ACTION_CONTINUOUS:
            shape, dt = (self.n, 2), np.float32
        elif self.cfg.action_mode == abi.ACTION_CONST_SPEED:
            shape, dt = (self.n, 1), np.float32
        else:
            shape, dt = (self.n,), np.int32
        self._actions = alloc(shape, dt) if alloc else np.zeros(shape, dt)
        self._pool = None
        self._stream = None
        if own_stream:
            self._stream = self._L.ftl_host_stream(self._h)
            if not self._stream:
                raise FtlError("ftl_host_stream failed")

    def close(self):
        if getattr(self, "_h", None):
            self._L.ftl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scenarios(self, pool):
        st = pool.c_struct()
        check(self._L, self._L.ftl_upload_scenarios(self._h, C.byref(st)), "ftl_upload_scenarios")
        self._pool = pool

    def reset(self, mask=None, scenario_ids=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        s = None if scenario_ids is None else np.ascontiguousarray(scenario_ids, np.int32)
        check(self._L, self._L.ftl_reset_host(self._h, abi.ptr(m), abi.ptr(s), C.byref(self.out.c), self._stream),
              "ftl_reset_host")
        return self.out

    def step(self, actions):
        np.copyto(self._actions, np.asarray(actions).reshape(self._actions.shape), casting="same_kind")
        check(self._L, self._L.ftl_step_host(self._h, abi.ptr(self._actions), C.byref(self.out.c), self._stream),
              "ftl_step_host")
        return self.out

    def step_begin(self, actions):
        """Enqueue one step (ftl_step_host_begin) and return at once; ``step_wait`` delivers the outputs."""
        np.copyto(self._actions, np.asarray(actions).reshape(self._actions.shape), casting="same_kind")
        check(self._L, self._L.ftl_step_host_begin(self._h, abi.ptr(self._actions), C.byref(self.out.c), self._stream),
              "ftl_step_host_begin")

    def step_wait(self):
        check(self._L, self._L.ftl_step_host_wait(self._h), "ftl_step_host_wait")
        return self.out

    def get_state(self, first=0, n=None):
        n = self.n - first if n is None else n
        st = HostState(n, self.cfg)
        check(self._L, self._L.ftl_get_state(self._h, first, n, C.byref(st.c)), "ftl_get_state")
        return st

    def set_state(self, st, first=0):
        check(self._L, self._L.ftl_set_state(self._h, first, len(st.env), C.byref(st.c)), "ftl_set_state")

    @property
    def h2d_bytes_per_step(self):
        return int(self._actions.nbytes)

    @property
    def d2h_bytes_per_step(self):
        return int(self.out.nbytes())
