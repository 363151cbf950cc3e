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

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
# every entry point include/ftl.h declares: name -> (argtypes, restype)
SIGNATURES = {
    "ftl_abi_version": ([], C.c_int),
    "ftl_last_error": ([], C.c_char_p),
    "ftl_build_info": ([], C.c_char_p),
    "ftl_create": ([C.POINTER(abi.FtlConfig), _i32, _i32, _i64, C.POINTER(_vp)], C.c_int),
    "ftl_destroy": ([_vp], C.c_int),
    "ftl_rays_per_env": ([_vp], C.c_int),
    "ftl_num_envs": ([_vp], C.c_int),
    "ftl_laser_beam_count": ([C.c_double, C.c_double], C.c_int),
    "ftl_upload_scenarios": ([_vp, C.POINTER(abi.FtlScenarioPool)], C.c_int),
    "ftl_reset": ([_vp, _vp, _vp, C.POINTER(abi.FtlOutputs), _vp], C.c_int),
    "ftl_step": ([_vp, _vp, C.POINTER(abi.FtlOutputs), _vp], C.c_int),
    "ftl_step_ex": ([_vp, _vp, C.POINTER(abi.FtlStepInputs), C.POINTER(abi.FtlOutputs), _vp], C.c_int),
    "ftl_step_host_ex": ([_vp, _vp, C.POINTER(abi.FtlStepInputs), C.POINTER(abi.FtlOutputs), _vp], C.c_int),
    "ftl_reset_host": ([_vp, _vp, _vp, C.POINTER(abi.FtlOutputs), _vp], C.c_int),
    "ftl_step_host": ([_vp, _vp, C.POINTER(abi.FtlOutputs), _vp], C.c_int),
    "ftl_step_host_begin": ([_vp, _vp, C.POINTER(abi.FtlOutputs), _vp], C.c_int),
    "ftl_step_host_wait": ([_vp], C.c_int),
    "ftl_host_stream": ([_vp], _vp),
    "ftl_get_state": ([_vp, _i32, _i32, C.POINTER(abi.FtlStateBuffers)], C.c_int),
    "ftl_set_state": ([_vp, _i32, _i32, C.POINTER(abi.FtlStateBuffers)], C.c_int),
    "ftl_stats": ([_vp, _vp, _i32, _vp], C.c_int),
    "ftl_set_option": ([_vp, _i32, _i32], C.c_int),
    "ftl_render": ([_vp, _i32, _i32, _i32, _vp, _vp], C.c_int),
    "ftl_render_host": ([_vp, _i32, _i32, _i32, _vp], C.c_int),
    "ftl_launch_count": ([_vp], _i64),
    "ftl_profile": ([_vp, _i32], C.c_int),
    "ftl_profile_read": ([_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)], C.c_int),
    "ftl_profile_read_kernels": ([_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)],
                                 C.c_int),
    "ftl_policy_mlp": ([C.POINTER(abi.FtlMlpWeights), _vp, _i32, _vp, _i32, _vp, _vp, _vp], C.c_int),
    "ftl_measure_fp32_peak": ([_i32, C.POINTER(C.c_double)], C.c_int),
    "ftl_generate_scenarios": ([C.POINTER(abi.FtlScenarioGenConfig), _vp, _i32, C.POINTER(abi.FtlScenarioPool), _i32],
                               C.c_int),
}


def bind(L, names):
    """Declare the ctypes signatures of `names` on the loaded library L (AttributeError if one is missing)."""
    for name in names:
        fn = getattr(L, name)
        fn.argtypes, fn.restype = SIGNATURES[name]
    return L


def load(path=None):
    """dlopen libftl.so (the CUDA build) and declare every include/ftl.h entry point.  A library that lacks one of
    them is refused: the product path has exactly one implementation."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise FtlLibraryMissing(
            "%s not found: build the CUDA extension first (python -m continiousenvironment_follower_leader_b200.build); "
            "there is no CPU fallback" % path)
    L = C.CDLL(path)
    try:
        bind(L, SIGNATURES)
    except AttributeError as e:
        raise FtlLibraryMissing("%s does not export the whole include/ftl.h C-ABI (%s)" % (path, e)) from None
    if L.ftl_abi_version() != abi.FTL_ABI_VERSION:
        raise FtlLibraryMissing("%s has ABI version %d, this package needs %d: rebuild it"
                                % (path, L.ftl_abi_version(), abi.FTL_ABI_VERSION))
    _LIBS[path] = L
    return L


def measure_fp32_peak(device=0):
    """FP32 FMA peak of `device` in TFLOP/s (ftl_measure_fp32_peak)."""
    L = load()
    v = C.c_double()
    check(L, L.ftl_measure_fp32_peak(int(device), C.byref(v)), "ftl_measure_fp32_peak")
    return v.value


def check(L, rc, what):
    if rc != 0:
        msg = L.ftl_last_error()
        msg = msg.decode() if msg else ""
        if rc == abi.FTL_ERR_INVALID:
            raise ValueError("%s: %s" % (what, msg))
        raise FtlError("%s failed (%d): %s" % (what, rc, msg))


class HostOutputs:
    """numpy buffers laid out as FtlOutputs."""

    def __init__(self, n, rays_per_env, alloc=None, follower_info=False, track_vector_len=0, radar_sectors=0,
                 laser_shape=None):
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
        self.laser = alloc((n,) + tuple(laser_shape), np.float32) if laser_shape else None   # LaserSensor, SEN:63-136
        self.c = abi.FtlOutputs(abi.ptr(self.numerical_features), abi.ptr(self.leader_target), abi.ptr(self.rays),
                                abi.ptr(self.reward), abi.ptr(self.done), abi.ptr(self.status),
                                None if self.follower_info is None else abi.ptr(self.follower_info),
                                None if self.track_vectors is None else abi.ptr(self.track_vectors),
                                None if self.radar is None else abi.ptr(self.radar),
                                None if self.laser is None else abi.ptr(self.laser))

    def nbytes(self):
        return sum(a.nbytes for a in (self.numerical_features, self.leader_target, self.rays, self.reward, self.done,
                                      self.status, self.follower_info, self.track_vectors, self.radar, self.laser) if a is not None)


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
                               track_vector_len=self.cfg.track_vector_len, radar_sectors=self.cfg.radar_sectors,
                               laser_shape=getattr(game_config, "laser_shape", None))
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

    def step(self, actions, frames=None, regime_draws=None):
        """One Game.step for every env.  Optional per-step inputs (FtlStepInputs, include/ftl.h): `frames` int32[N] =
        frames this step runs per env (random_frames_per_step, ENV:939-940); `regime_draws` float64[N, frames_per_step]
        = the random() behind random.uniform of list-valued leader_speed_regime entries (ENV:1155-1156)."""
        np.copyto(self._actions, np.asarray(actions).reshape(self._actions.shape), casting="same_kind")
        if frames is None and regime_draws is None:
            check(self._L, self._L.ftl_step_host(self._h, abi.ptr(self._actions), C.byref(self.out.c), self._stream),
                  "ftl_step_host")
            return self.out
        f = None if frames is None else np.ascontiguousarray(np.broadcast_to(frames, (self.n,)), np.int32)
        d = None if regime_draws is None else np.ascontiguousarray(regime_draws, np.float64)
        if d is not None and d.shape != (self.n, self.cfg.frames_per_step):
            raise ValueError("regime_draws must have shape (n_envs, frames_per_step)")
        ins = abi.FtlStepInputs(abi.ptr(f), abi.ptr(d))
        check(self._L, self._L.ftl_step_host_ex(self._h, abi.ptr(self._actions), C.byref(ins), C.byref(self.out.c),
                                                self._stream), "ftl_step_host_ex")
        return self.out

    def step_begin(self, actions):
        """Enqueue one step (ftl_step_host_begin) and return at once; ``step_wait`` delivers the outputs."""
        np.copyto(self._actions, np.asarray(actions).reshape(self._actions.shape), casting="same_kind")
        check(self._L, self._L.ftl_step_host_begin(self._h, abi.ptr(self._actions), C.byref(self.out.c), self._stream),
              "ftl_step_host_begin")

    def step_wait(self):
        check(self._L, self._L.ftl_step_host_wait(self._h), "ftl_step_host_wait")
        return self.out

    def render(self, first=0, n=None, scale=1):
        """rgb_array of envs [first, first + n): uint8 [n, H, W, 3], H = ceil(game_height / scale) (ftl_render_host)."""
        n = self.n - first if n is None else n
        W, H = -(-self.cfg.game_width // scale), -(-self.cfg.game_height // scale)
        img = np.zeros((n, H, W, 3), np.uint8)
        check(self._L, self._L.ftl_render_host(self._h, int(first), int(n), int(scale), abi.ptr(img)), "ftl_render_host")
        return img

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
