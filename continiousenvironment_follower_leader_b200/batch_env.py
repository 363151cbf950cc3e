"""Device-resident batched environment: N reference ``Game`` instances advanced per kernel launch.

PyTorch is plumbing here (device memory, streams, torch.distributed); the step itself is libftl.so.
``FtlBatchEnv.step(actions)`` mirrors ``Game.step`` (follow_the_leader_continuous_env.py:908-945) for a
batch: ``actions`` is a CUDA tensor [N, 2] (or [N, 1] / int32 [N] for the constant-speed / discrete
action spaces), the return value is ``(obs, reward, done, info)`` with CUDA tensors:

    obs["numerical_features"]  float32 [N, 10]                      ENV:1793-1802
    obs["leader_target_point"] int32   [N, 2]                       ENV:1803-1806
    obs[<ray sensor name>]     float32 [N, H, R] (or [N, H, 4R])    SEN:883-962
    obs[<flat ray sensor>]     float32 [N, R]                       LeaderCorridor_lasers(_v2), SEN:571-807
    obs[<FollowerInfo>]        float32 [N, 2]                       SEN:834-842
    obs[<LeaderTrackDetector_vector>] float32 [N, P, 2]            SEN:365-380
    obs[<LeaderTrackDetector_radar>]  float32 [N, sectors]         SEN:425-461
    obs[<LaserSensor>]         float32 [N, beams, 2] (or [N, beams])   SEN:63-136
    obs["sensor_prev"]         float32 [N, H, sum of widths]        WRP:203-221, instead of the sensor entries when the
                                                                    config was built with fused_sensor_prev=True
    reward float32 [N], done bool [N]
    info["status"] uint8 [N, 4] = (mission_status, agent_status, leader_status, crash) codes, see abi.py
"""
import ctypes as C

import numpy as np
import torch

from . import abi, capi
from .config import GameConfig


class FtlBatchEnv:
    def __init__(self, n_envs, game_config=None, scenario_pool=None, device=None, env_id_base=0, lib_path=None,
                 **game_kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("FtlBatchEnv needs a CUDA device (there is no CPU fallback)")
        self.gc = game_config if game_config is not None else GameConfig(**game_kwargs)
        self.cfg = self.gc.c
        self.n = int(n_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._L = capi.load(lib_path)
        self._h = C.c_void_p()
        capi.check(self._L, self._L.ftl_create(C.byref(self.cfg), self.n, self.device.index or 0, int(env_id_base),
                                               C.byref(self._h)), "ftl_create")
        dev, n, rpe = self.device, self.n, self.gc.rays_per_env
        self.numerical_features = torch.zeros((n, 10), dtype=torch.float32, device=dev)
        self.leader_target = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        self.rays = torch.zeros((n, max(rpe, 1)), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.status = torch.zeros((n, 4), dtype=torch.uint8, device=dev)
        # optional sensors: FollowerInfo [N, 2], LeaderTrackDetector_vector [N, P, 2]
        self.follower_info = torch.zeros((n, 2), dtype=torch.float32, device=dev) \
            if self.gc.follower_info_name is not None else None
        self.track_vectors = torch.zeros((n, self.cfg.track_vector_len, 2), dtype=torch.float32, device=dev) \
            if self.cfg.track_vector_len else None
        self.radar = torch.zeros((n, self.cfg.radar_sectors), dtype=torch.float32, device=dev) \
            if self.cfg.radar_sectors else None
        self.laser = torch.zeros((n,) + tuple(self.gc.laser_shape), dtype=torch.float32, device=dev) \
            if self.gc.laser_shape else None
        self._out = abi.FtlOutputs(self.numerical_features.data_ptr(), self.leader_target.data_ptr(),
                                   self.rays.data_ptr() if rpe else None, self.reward.data_ptr(),
                                   self.done.data_ptr(), self.status.data_ptr(),
                                   None if self.follower_info is None else self.follower_info.data_ptr(),
                                   None if self.track_vectors is None else self.track_vectors.data_ptr(),
                                   None if self.radar is None else self.radar.data_ptr(),
                                   None if self.laser is None else self.laser.data_ptr())
        self._stats = torch.zeros(abi.STAT_COUNT, dtype=torch.float64, device=dev)
        self._ray_layout = self.gc.ray_layout()
        self._env_id_base = int(env_id_base)
        self._frames_gen = None
        self.last_frames = None
        self._pool = None
        if scenario_pool is not None:
            self.upload_scenarios(scenario_pool)

    # ---- lifecycle -----------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            torch.cuda.synchronize(self.device)
            self._L.ftl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scenarios(self, pool):
        st = pool.c_struct()
        capi.check(self._L, self._L.ftl_upload_scenarios(self._h, C.byref(st)), "ftl_upload_scenarios")
        self._pool = pool

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- gym-like surface ----------------------------------------------------------------------------
    def _obs(self):
        obs = {"numerical_features": self.numerical_features, "leader_target_point": self.leader_target}
        if self.follower_info is not None:
            obs[self.gc.follower_info_name] = self.follower_info
        if self.track_vectors is not None:
            obs[self.gc.track_vector_name] = self.track_vectors
        if self.radar is not None:
            obs[self.gc.radar_name] = self.radar
        if self.laser is not None:
            obs[self.gc.laser_name] = self.laser
        if self.cfg.fused_sensor_prev:   # one [N, H, sum of widths] matrix, already clip(v / laser_length, 0, 1)
            obs["sensor_prev"] = self.sensor_prev()
            return obs
        for k, (name, off, h, w) in enumerate(self._ray_layout):
            block = self.rays[:, off:off + h * w]
            obs[name] = block if self.gc.ray_sensor_flat[k] else block.view(self.n, h, w)
        return obs

    def sensor_prev(self):
        """What ContinuousObserveModifier_sensorPrev.observation returns (WRP:203-221), batched: [N, H, sum W]."""
        if not self.cfg.fused_sensor_prev:
            from .wrappers import sensor_prev_observation
            return sensor_prev_observation(self)
        h = self._ray_layout[0][2]
        return self.rays.view(self.n, h, self.gc.rays_per_env // h)

    def reset(self, mask=None, scenario_ids=None):
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        s = None if scenario_ids is None else scenario_ids.to(device=self.device, dtype=torch.int32).contiguous()
        capi.check(self._L, self._L.ftl_reset(self._h, None if m is None else m.data_ptr(),
                                              None if s is None else s.data_ptr(), C.byref(self._out), self._stream()),
                   "ftl_reset")
        return self._obs()

    def step(self, actions, frames=None, regime_draws=None):
        """frames: int32 CUDA tensor [N], frames this step runs per env (random_frames_per_step, ENV:939-940; drawn
        on the device from `random_frames_per_step` when the configuration has it and none is given); regime_draws:
        float64 CUDA tensor [N, frames_per_step], see FtlStepInputs in include/ftl.h."""
        want = torch.int32 if self.cfg.action_mode == abi.ACTION_DISCRETE else torch.float32
        if actions.dtype != want or not actions.is_contiguous() or actions.device != self.device:
            actions = actions.to(device=self.device, dtype=want).contiguous()
        if frames is None and self.gc.random_frames_per_step is not None:
            lo, hi = self.gc.random_frames_per_step     # np.random.randint(lo, hi): hi exclusive
            if self._frames_gen is None:
                self._frames_gen = torch.Generator(device=self.device).manual_seed(0x5eed + int(self._env_id_base))
            frames = torch.randint(int(lo), int(hi), (self.n,), generator=self._frames_gen, device=self.device,
                                   dtype=torch.int32)
        if frames is None and regime_draws is None:
            capi.check(self._L, self._L.ftl_step(self._h, actions.data_ptr(), C.byref(self._out), self._stream()), "ftl_step")
        else:
            f = None if frames is None else frames.to(device=self.device, dtype=torch.int32).contiguous()
            d = None if regime_draws is None else regime_draws.to(device=self.device, dtype=torch.float64).contiguous()
            if d is not None and tuple(d.shape) != (self.n, self.cfg.frames_per_step):
                raise ValueError("regime_draws must have shape (n_envs, frames_per_step)")
            ins = abi.FtlStepInputs(None if f is None else f.data_ptr(), None if d is None else d.data_ptr())
            capi.check(self._L, self._L.ftl_step_ex(self._h, actions.data_ptr(), C.byref(ins), C.byref(self._out),
                                                    self._stream()), "ftl_step_ex")
            self.last_frames = f
        return self._obs(), self.reward, self.done.bool(), {"status": self.status}

    def step_raw(self, actions):
        """step() without building the python-side views (what bench.py times)."""
        capi.check(self._L, self._L.ftl_step(self._h, actions.data_ptr(), C.byref(self._out), self._stream()), "ftl_step")

    def render(self, first=0, n=1, scale=4):
        """rgb_array frames of envs [first, first + n) as a uint8 device tensor [n, H, W, 3] (ftl_render, ENV:1196-1302;
        H = ceil(game_height / scale)): a debugging / video view drawn from the state on the device."""
        W, H = -(-self.cfg.game_width // scale), -(-self.cfg.game_height // scale)
        img = torch.empty((n, H, W, 3), dtype=torch.uint8, device=self.device)
        capi.check(self._L, self._L.ftl_render(self._h, int(first), int(n), int(scale), img.data_ptr(), self._stream()),
                   "ftl_render")
        return img

    # ---- state / statistics ----------------------------------------------------------------------------
    def get_state(self, first=0, n=None):
        n = self.n - first if n is None else n
        st = capi.HostState(n, self.cfg)
        capi.check(self._L, self._L.ftl_get_state(self._h, first, n, C.byref(st.c)), "ftl_get_state")
        return st

    def set_state(self, st, first=0):
        capi.check(self._L, self._L.ftl_set_state(self._h, first, len(st.env), C.byref(st.c)), "ftl_set_state")

    def stats(self, reset=False, reduce_across_ranks=False):
        """Episode statistics summed over this handle's envs (and over all ranks with NCCL if asked)."""
        capi.check(self._L, self._L.ftl_stats(self._h, self._stats.data_ptr(), int(bool(reset)), self._stream()), "ftl_stats")
        out = self._stats.clone()
        if reduce_across_ranks and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(out, op=torch.distributed.ReduceOp.SUM)
        return out

    def stats_dict(self, **kw):
        v = self.stats(**kw).cpu().numpy()
        return {name: float(v[i]) for i, name in enumerate(abi.STAT_NAMES)}

    def profile(self, enable=True):
        capi.check(self._L, self._L.ftl_profile(self._h, int(bool(enable))), "ftl_profile")

    def profile_read(self):
        """(ms in the fused step kernel, ms in the ray kernel, steps) accumulated since profile(True)."""
        a, b, k = C.c_double(), C.c_double(), C.c_int64()
        capi.check(self._L, self._L.ftl_profile_read(self._h, C.byref(a), C.byref(b), C.byref(k)), "ftl_profile_read")
        return a.value, b.value, k.value

    def profile_read_kernels(self):
        """{"k_kin", "k_book", "k_rays"} ms accumulated since profile(True) (k_rays includes k_finish), and the steps."""
        a, b, r, k = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        capi.check(self._L, self._L.ftl_profile_read_kernels(self._h, C.byref(a), C.byref(b), C.byref(r), C.byref(k)),
                   "ftl_profile_read_kernels")
        return {"k_kin": a.value, "k_book": b.value, "k_rays": r.value}, k.value

    @property
    def launch_count(self):
        return int(self._L.ftl_launch_count(self._h))


def info_strings(status_row):
    """uint8[4] status codes -> the reference's info dict (ENV:951-955)."""
    return {"mission_status": abi.MISSION_STATUS[int(status_row[0])], "agent_status": abi.AGENT_STATUS[int(status_row[1])],
            "leader_status": abi.LEADER_STATUS[int(status_row[2])]}
