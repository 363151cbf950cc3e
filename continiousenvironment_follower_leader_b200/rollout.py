"""Device-resident consumer of the simulator: a rollout loop that never leaves the GPU.

The reference trains through RLlib (notebooks/Ray_train_demo.ipynb): one ``Game`` per rollout worker wrapped in
``ContinuousObserveModifier_sensorPrev`` (utils/wrappers.py:169-221), observations pickled to the learner.  Here the
wrapper's matrix is written by the ray kernel itself (``fused_sensor_prev``) STRAIGHT INTO the trajectory ring (row t + 1
of the observation archive is the output buffer of step t), the policy reads row t in place and writes its action into
the row ``ftl_step`` consumes, and ``ftl_step`` writes reward and done into their rings -- no copy of any kind per step,
one stream, no synchronisation inside ``collect``.  The built-in ``MlpPolicy`` runs as ONE fused tensor-core kernel of
libftl.so (``ftl_policy_mlp``, csrc/ftl_policy.cu: bfloat16 ``mma``, activations in shared memory, the action written
where ``ftl_step`` reads it), so a step costs two launches; any other torch module runs through torch, its section of a
step captured in one CUDA graph per ring slot.

PyTorch is the consumer here (a user's policy network is a torch module; its GEMMs are library calls); the simulator
side is libftl.so exactly as in ``FtlBatchEnv.step``.
"""
import ctypes as C

import torch

from . import abi, capi
from .batch_env import FtlBatchEnv
from .config import GameConfig


class MlpPolicy(torch.nn.Module):
    """tanh-squashed Gaussian policy + value head on the flattened sensorPrev matrix (the shape the shipped trained
    configurations feed their fully connected nets).  ``forward`` returns (action, value); the mean and the value come
    out of one fused head."""

    def __init__(self, obs_dim, act_low, act_high, hidden=128, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)

        def lin(i, o):
            layer = torch.nn.Linear(i, o)
            with torch.no_grad():
                layer.weight.copy_(torch.randn(o, i, generator=g) / i ** 0.5)
                layer.bias.zero_()
            return layer

        self.act_dim = len(act_low)
        self.body = torch.nn.Sequential(lin(obs_dim, hidden), torch.nn.Tanh(), lin(hidden, hidden), torch.nn.Tanh())
        self.head = lin(hidden, self.act_dim + 1)          # [mu..., value]
        self.log_std = torch.nn.Parameter(torch.full((self.act_dim,), -0.5))
        self.register_buffer("act_mid", torch.as_tensor((act_high + act_low) / 2, dtype=torch.float32))
        self.register_buffer("act_half", torch.as_tensor((act_high - act_low) / 2, dtype=torch.float32))

    def forward(self, obs_flat, noise=None):
        out = self.head(self.body(obs_flat))
        mu, value = out[:, :self.act_dim], out[:, self.act_dim]
        if noise is not None:
            mu = mu + noise * self.log_std.exp()
        return self.act_mid + self.act_half * torch.tanh(mu), value


class DeviceRollout:
    """``collect(T)`` advances all N envs T steps under ``policy`` and returns the trajectory as device tensors.

    Rings (allocated once): obs [T+1, N, H*W] float32 (the simulator's own output rows), actions [T, N, A], rewards
    [T, N], dones [T, N] (uint8), values [T+1, N].  ``use_graphs``: capture the policy section per ring slot in a CUDA
    graph.
    """

    def __init__(self, n_envs, horizon, game_config=None, scenario_pool=None, policy=None, device=None, env_id_base=0,
                 seed=0, use_graphs=True, allow_tf32=True, fused_policy=True, **game_kwargs):
        if game_config is None:
            game_config = GameConfig(fused_sensor_prev=True, auto_reset=True, **game_kwargs)
        if not game_config.c.fused_sensor_prev or not game_config.c.auto_reset:
            raise ValueError("DeviceRollout needs a GameConfig built with fused_sensor_prev=True and auto_reset=True")
        if game_config.discrete_action_space:
            raise NotImplementedError("DeviceRollout drives the continuous action spaces")
        self.env = FtlBatchEnv(n_envs, game_config=game_config, scenario_pool=scenario_pool, device=device,
                               env_id_base=env_id_base)
        self.n, self.T = int(n_envs), int(horizon)
        dev = self.env.device
        self.obs_dim = game_config.rays_per_env
        lo, hi = game_config.action_bounds()
        self.act_dim = len(lo)
        self.policy = policy if policy is not None else MlpPolicy(self.obs_dim, lo, hi, seed=seed)
        self.policy.to(dev)
        self.allow_tf32 = bool(allow_tf32)
        self.obs = torch.zeros((self.T + 1, self.n, self.obs_dim), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((self.T, self.n, self.act_dim), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((self.T, self.n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((self.T, self.n), dtype=torch.uint8, device=dev)
        self.values = torch.zeros((self.T + 1, self.n), dtype=torch.float32, device=dev)
        self._noise = torch.zeros((self.n, self.act_dim), dtype=torch.float32, device=dev)
        # one FtlOutputs per ring slot: the kernels of step t write the next observation into obs[t + 1] and reward /
        # done into row t of their rings
        env = self.env
        self._outs = [abi.FtlOutputs(env.numerical_features.data_ptr(), env.leader_target.data_ptr(),
                                     self.obs[t + 1].data_ptr(), self.rewards[t].data_ptr(), self.dones[t].data_ptr(),
                                     env.status.data_ptr(), None, None, None, None) for t in range(self.T)]
        self._out_reset = abi.FtlOutputs(env.numerical_features.data_ptr(), env.leader_target.data_ptr(),
                                         self.obs[0].data_ptr(), env.reward.data_ptr(), env.done.data_ptr(),
                                         env.status.data_ptr(), None, None, None, None)
        self._last = 0     # ring row that holds the current observation
        self._fused = None
        if fused_policy and isinstance(self.policy, MlpPolicy) and self.policy.head.in_features == 128 \
                and self.obs_dim % 16 == 0 and self.obs_dim <= 288 and self.act_dim <= 7:
            self._noise_ring = torch.zeros((self.T, self.n, self.act_dim), dtype=torch.float32, device=dev)
            self._scratch_act = torch.zeros((self.n, self.act_dim), dtype=torch.float32, device=dev)
            self.sync_policy()
        self._graphs = {}
        self._use_graphs = bool(use_graphs)
        self._seed = int(seed)
        self._started = False

    def close(self):
        self._graphs.clear()
        self.env.close()

    def sync_policy(self):
        """(Re-)export the MlpPolicy's parameters for the fused kernel: bfloat16 weights in torch.nn.Linear layout,
        float32 biases.  Call after every optimiser step on the policy."""
        p = self.policy
        l1, l2 = p.body[0], p.body[2]
        keep = {"w1": l1.weight.detach().to(torch.bfloat16).contiguous(), "b1": l1.bias.detach().float().contiguous(),
                "w2": l2.weight.detach().to(torch.bfloat16).contiguous(), "b2": l2.bias.detach().float().contiguous(),
                "w3": p.head.weight.detach().to(torch.bfloat16).contiguous(), "b3": p.head.bias.detach().float().contiguous(),
                "noise_scale": p.log_std.detach().float().exp().contiguous(),
                "act_mid": p.act_mid.float().contiguous(), "act_half": p.act_half.float().contiguous()}
        self._fused = (abi.FtlMlpWeights(*[keep[k].data_ptr() for k in ("w1", "b1", "w2", "b2", "w3", "b3", "noise_scale",
                                                                         "act_mid", "act_half")],
                                         self.obs_dim, self.act_dim), keep)

    def _fused_policy(self, obs_row, noise_row, actions_row, values_row, stream):
        env = self.env
        capi.check(env._L, env._L.ftl_policy_mlp(C.byref(self._fused[0]), obs_row.data_ptr(), self.obs_dim,
                                                 None if noise_row is None else noise_row.data_ptr(), self.n,
                                                 actions_row.data_ptr(), values_row.data_ptr(), stream), "ftl_policy_mlp")

    def _policy_section(self, t, explore):
        """run the policy on obs[t] in place, leave the action in actions[t] (which ftl_step reads) and the value in
        values[t]"""
        noise = self._noise.normal_() if explore else None
        act, val = self.policy(self.obs[t], noise)
        self.actions[t].copy_(act)
        self.values[t].copy_(val)

    def _run_policy(self, t, explore):
        if not self._use_graphs:
            return self._policy_section(t, explore)
        key = (t, bool(explore))
        g = self._graphs.get(key)
        if g is None:
            # warm up on a side stream (library handles, autotuning), then capture this slot's section
            side = torch.cuda.Stream(device=self.env.device)
            side.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._policy_section(t, explore)
            torch.cuda.current_stream(self.env.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._policy_section(t, explore)
            self._graphs[key] = g
        g.replay()

    @torch.no_grad()
    def collect(self, steps=None, explore=True):
        """T steps (default: the horizon); everything is enqueued on the current stream, nothing is synchronised."""
        T = self.T if steps is None else int(steps)
        if T > self.T:
            raise ValueError("steps exceeds the horizon the buffers were allocated for")
        env = self.env
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.allow_tf32
        try:
            stream = env._stream()
            if not self._started:
                torch.cuda.manual_seed(self._seed)
                capi.check(env._L, env._L.ftl_reset(env._h, None, None, C.byref(self._out_reset), stream), "ftl_reset")
                self._started = True
            elif self._last != 0:
                self.obs[0].copy_(self.obs[self._last])     # once per collect: the window starts where the last one ended
            if self._fused is not None and explore:
                self._noise_ring[:T].normal_()          # one kernel per collect
            for t in range(T):
                if self._fused is not None:
                    self._fused_policy(self.obs[t], self._noise_ring[t] if explore else None, self.actions[t],
                                       self.values[t], stream)
                else:
                    self._run_policy(t, explore)
                capi.check(env._L, env._L.ftl_step(env._h, self.actions[t].data_ptr(), C.byref(self._outs[t]), stream),
                           "ftl_step")
            self._last = T
            if self._fused is not None:
                self._fused_policy(self.obs[T], None, self._scratch_act, self.values[T], stream)
            else:
                _, val = self.policy(self.obs[T], None)
                self.values[T].copy_(val)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        return {"obs": self.obs[:T + 1], "actions": self.actions[:T],
                "rewards": self.rewards[:T], "dones": self.dones[:T], "values": self.values[:T + 1]}

    @torch.no_grad()
    def advantages(self, steps=None, gamma=0.99, lam=0.95):
        """Generalised advantage estimation over the collected window, on the device (auto-reset: an env's ``done`` row
        cuts the bootstrap, the next row already belongs to the new episode)."""
        T = self.T if steps is None else int(steps)
        adv = torch.zeros((T, self.n), dtype=torch.float32, device=self.env.device)
        last = torch.zeros(self.n, dtype=torch.float32, device=self.env.device)
        for t in range(T - 1, -1, -1):
            live = 1.0 - self.dones[t].float()
            delta = self.rewards[t] + gamma * self.values[t + 1] * live - self.values[t]
            last = delta + gamma * lam * live * last
            adv[t] = last
        return adv, adv + self.values[:T]
