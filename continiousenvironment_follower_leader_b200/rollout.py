"""Device-resident consumer of the simulator: a rollout loop that never leaves the GPU.

The reference trains through RLlib (notebooks/Ray_train_demo.ipynb): one ``Game`` per rollout worker wrapped in
``ContinuousObserveModifier_sensorPrev`` (utils/wrappers.py:169-221), observations pickled to the learner.  Here the
wrapper's matrix is written by the ray kernel itself (``fused_sensor_prev``), a policy reads it IN PLACE from the env's
output buffer, writes its action straight into the tensor ``ftl_step`` consumes, and the trajectory lands in
pre-allocated device rings -- no host copy per step, one stream, no synchronisation inside ``collect``.

PyTorch is the consumer here (a user's policy network is a torch module; its GEMMs are library calls); the simulator
side is libftl.so exactly as in ``FtlBatchEnv.step``.
"""
import torch

from .batch_env import FtlBatchEnv
from .config import GameConfig


class MlpPolicy(torch.nn.Module):
    """tanh-squashed Gaussian policy + value head on the flattened sensorPrev matrix (the shape the shipped trained
    configurations feed their fully connected nets)."""

    def __init__(self, obs_dim, act_low, act_high, hidden=128, seed=0, dtype=torch.float32):
        super().__init__()
        g = torch.Generator().manual_seed(seed)

        def lin(i, o):
            layer = torch.nn.Linear(i, o)
            with torch.no_grad():
                layer.weight.copy_(torch.randn(o, i, generator=g) / i ** 0.5)
                layer.bias.zero_()
            return layer

        self.body = torch.nn.Sequential(lin(obs_dim, hidden), torch.nn.Tanh(), lin(hidden, hidden), torch.nn.Tanh())
        self.mu = lin(hidden, len(act_low))
        self.value = lin(hidden, 1)
        self.log_std = torch.nn.Parameter(torch.full((len(act_low),), -0.5))
        self.register_buffer("act_mid", torch.as_tensor((act_high + act_low) / 2, dtype=torch.float32))
        self.register_buffer("act_half", torch.as_tensor((act_high - act_low) / 2, dtype=torch.float32))
        self.to(dtype)

    def forward(self, obs_flat, noise=None):
        h = self.body(obs_flat)
        mu = self.mu(h)
        if noise is not None:
            mu = mu + noise * self.log_std.exp()
        return self.act_mid + self.act_half * torch.tanh(mu.float()), self.value(h).squeeze(-1).float()


class DeviceRollout:
    """``collect(T)`` advances all N envs T steps under ``policy`` and returns the trajectory as device tensors.

    Buffers (allocated once): obs [T+1, N, H*W] (``obs_dtype``), actions [T, N, A], rewards [T, N], dones [T, N],
    values [T+1, N].  With ``store_obs=False`` only the newest observation is kept (evaluation / throughput runs).
    """

    def __init__(self, n_envs, horizon, game_config=None, scenario_pool=None, policy=None, device=None, env_id_base=0,
                 obs_dtype=torch.float32, store_obs=True, seed=0, **game_kwargs):
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
        self.store_obs = bool(store_obs)
        t_obs = self.T + 1 if self.store_obs else 1
        self.obs = torch.zeros((t_obs, self.n, self.obs_dim), dtype=obs_dtype, device=dev)
        self.actions = torch.zeros((self.T, self.n, self.act_dim), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((self.T, self.n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((self.T, self.n), dtype=torch.uint8, device=dev)
        self.values = torch.zeros((self.T + 1, self.n), dtype=torch.float32, device=dev)
        self._gen = torch.Generator(device=dev).manual_seed(seed)
        self._noise = torch.zeros((self.n, self.act_dim), dtype=torch.float32, device=dev)
        self._started = False

    def close(self):
        self.env.close()

    def _current_obs(self):
        # the ray kernel's output buffer viewed as [N, H*W]: no copy
        return self.env.rays.view(self.n, self.obs_dim)

    @torch.no_grad()
    def collect(self, steps=None, explore=True):
        """T steps (default: the horizon); everything is enqueued on the current stream, nothing is synchronised."""
        T = self.T if steps is None else int(steps)
        if T > self.T:
            raise ValueError("steps exceeds the horizon the buffers were allocated for")
        env = self.env
        if not self._started:
            env.reset()
            self._started = True
        w_dtype = next(self.policy.parameters()).dtype
        for t in range(T):
            cur = self._current_obs()
            if self.store_obs:
                self.obs[t].copy_(cur)           # the next ftl_step overwrites the env's buffer
            noise = None
            if explore:
                noise = self._noise.normal_(generator=self._gen)
            act, val = self.policy(cur if cur.dtype == w_dtype else cur.to(w_dtype), noise)
            self.actions[t].copy_(act)           # ftl_step reads this row in place
            self.values[t].copy_(val)
            env.step_raw(self.actions[t])
            self.rewards[t].copy_(env.reward)
            self.dones[t].copy_(env.done)
        cur = self._current_obs()
        if self.store_obs:
            self.obs[T].copy_(cur)
        _, val = self.policy(cur if cur.dtype == w_dtype else cur.to(w_dtype), None)
        self.values[T].copy_(val)
        return {"obs": self.obs[:T + 1] if self.store_obs else self.obs, "actions": self.actions[:T],
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
