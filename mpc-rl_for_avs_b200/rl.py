"""SURVEY 8-f "next" rows N1 + N2: a batched synthetic intersection environment and the batched
A2C-MPC / PPO-MPC rollout and update loops (BASELINE config 4: 1024 vectorised envs, RL-set reference
speed, MPC on the GPU).  NOT part of the hot path; plain PyTorch tensor code around `BatchedPureMPC`.

What is mirrored from the reference (file:line in SaeedRahmani/MPC-RL_for_AVs):
  * rollout structure of `A2C_MPC.collect_rollouts` (agents/a2c_mpc.py:111-180): policy(obs) -> RL action =
    reference speed (v0) -> MPC -> env.step(raw (a, delta)) -> buffer stores the *RL* action;
  * A2C hyper-parameters of config/cfg.yaml:30-45 (n_steps 64, lr 7e-4, RMSprop eps 1e-5, gamma 0.99,
    gae_lambda 1.0, vf_coef 0.5, max_grad_norm 0.5, ent_coef 0) and the loss of agents/a2c_mpc.py:182-226;
  * the policy's action space Box(-1, 1, (1,)) (trainers/trainer_utils.py:6-26), unclipped Gaussian samples
    as A2C passes them -> clip(., 0, 30) inside update_reference_states;
  * observation layout / ordering (config/config.py:10-26: Kinematics, absolute, sorted by distance);
  * reward and termination rules of envs/intersection_env__.py:61-129 (collision -5, speed reward over
    [7, 9] m/s, arrived +1, episode ends on crash / arrival / time limit);
  * the action path of the SB3 subclasses: raw (a, delta) into highway-env's ContinuousAction, i.e. clipped
    to [-1, 1] and scaled to +-5 m/s^2 / +-pi/4 (quirk Q8).

What is a STAND-IN (highway-env is not installable here): other vehicles drive straight at constant speed
on the four approach lanes and are respawned when they leave the map; collisions are a 2.5 m centre distance
test; "arrived" is reaching the end of the hard-coded path.
"""
from __future__ import annotations

import math
import time
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .scenarios import reference_path

_LANE_HEADING = (-math.pi / 2, 0.0, math.pi / 2, math.pi)


def _i64(x: int) -> int:
    """Python int -> the same 64-bit pattern as a signed value (torch int64 arithmetic wraps)."""
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


class BatchedIntersectionEnv:
    """B independent intersection scenes as tensors on one device (CPU works too: used by the CPU tests).

    `step` / `reset(mask)` / `observe` are fixed-shape tensor programs without any host synchronisation
    (no `.any()` / `.item()` branches, masked updates instead), and the random draws come from a counter-based
    generator kept on the device (splitmix64 over (seed, draw counter, env, slot)), so a whole
    policy -> MPC -> env transition can be captured in a CUDA graph (`_MPCRollout(graph=True)`)."""

    def __init__(self, n_envs: int, n_others: int = 9, device="cuda", seed: int = 0, policy_frequency: int = 10,
                 simulation_frequency: int = 30, duration_steps: int = 100, action_scaling: str = "sb3_raw",
                 fused: Optional[bool] = None):
        if action_scaling not in ("sb3_raw", "physical"):
            raise ValueError("action_scaling must be 'sb3_raw' or 'physical'")
        self.B, self.M, self.V = int(n_envs), int(n_others), int(n_others) + 1
        self.device = torch.device(device)
        self.sub = simulation_frequency // policy_frequency
        self.dt_sim = 1.0 / simulation_frequency
        self.duration_steps = duration_steps
        self.action_scaling = action_scaling
        ref = torch.from_numpy(reference_path(1.0 / policy_frequency)).to(self.device)
        self.ref_xy = ref[:, :2].float()
        self._arrive = (float(ref[-2, 0].float()), float(ref[-1, 1].float()))      # arrival test: x <= ref[-2].x, |y - ref[-1].y| < 4
        f = dict(device=self.device, dtype=torch.float32)
        self.ego = torch.zeros(self.B, 4, **f)               # x, y, heading, speed
        self.oth = torch.zeros(self.B, self.M, 4, **f)       # x, y, speed, heading
        self.t = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        self.crashed = torch.zeros(self.B, dtype=torch.bool, device=self.device)
        self._lane_heading = torch.tensor(_LANE_HEADING, **f)
        self._ctr = torch.full((1,), _i64(0x9E3779B97F4A7C15 * (2 * int(seed) + 1)), dtype=torch.int64, device=self.device)
        self._slot = torch.arange(self.B * 64, dtype=torch.int64, device=self.device).reshape(self.B, 64)
        # fused=True: `step` is one CUDA kernel of libmpcb200 (csrc/mpc_env.cu: same rules, same random numbers);
        # default on CUDA devices with <= 15 other vehicles.  The tensor program below is the specification and
        # the CPU path.
        self.fused = (self.device.type == "cuda" and self.M <= 15) if fused is None else bool(fused)
        if self.fused:
            if self.device.type != "cuda" or self.M > 15:
                raise ValueError("the fused step needs a CUDA device and at most 15 other vehicles")
            from . import _capi
            self._capi, self._lib = _capi, _capi.load()
        self.reset()

    # ---------------------------------------------------------------------------------------------- random draws
    def _rand(self, n: int) -> torch.Tensor:
        """[B, n] uniforms in (0, 1): splitmix64 of (draw counter, env * 64 + column); n <= 64."""
        self._ctr += _i64(0xD1B54A32D192ED03)
        x = self._slot[:, :n] * _i64(0x9E3779B97F4A7C15) + self._ctr
        x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * _i64(0xBF58476D1CE4E5B9)
        x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * _i64(0x94D049BB133111EB)
        x = x ^ ((x >> 31) & ((1 << 33) - 1))
        return (((x >> 40) & ((1 << 24) - 1)).float() + 0.5) * (1.0 / (1 << 24))

    def _randn(self, n: int) -> torch.Tensor:
        u = self._rand(2 * n)
        return torch.sqrt(-2.0 * torch.log(u[:, :n])) * torch.cos((2.0 * math.pi) * u[:, n:])

    def _spawn_others(self) -> torch.Tensor:
        """[B, M, 4] fresh vehicles on the four approach lanes (envs/intersection_env__.py:164-170, :409-416)."""
        M = self.M
        c = torch.clamp((self._rand(M) * 4.0).floor(), max=3.0)
        lane = 2.0 + 0.2 * self._randn(M)
        d = -30.0 + 100.0 * self._rand(M)
        ang = c * (math.pi / 2)
        px = torch.cos(ang) * lane - torch.sin(ang) * d
        py = torch.sin(ang) * lane + torch.cos(ang) * d
        spd = torch.clamp(8.0 + self._randn(M), min=0.5)
        hd = self._lane_heading[c.long()]
        return torch.stack([px, py, spd, hd], dim=-1)

    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        m = torch.ones(self.B, dtype=torch.bool, device=self.device) if mask is None else mask.to(self.device).bool()
        r = self._rand(2)
        ego = torch.stack([2.0 + 0.1 * self._randn(1)[:, 0],
                           50.0 + r[:, 0],                   # envs/intersection_env__.py:303-315: spawns at the south entry
                           torch.full((self.B,), -math.pi / 2, device=self.device),
                           8.0 + 2.0 * r[:, 1]], dim=1)
        oth = self._spawn_others()
        near = torch.hypot(oth[..., 0] - ego[:, None, 0], oth[..., 1] - ego[:, None, 1]) < 8.0
        oy = torch.where(near & (oth[..., 3] == -math.pi / 2), oth[..., 1] - 20.0, oth[..., 1])
        oth = torch.stack([oth[..., 0], oy, oth[..., 2], oth[..., 3]], dim=-1)
        # state lives in fixed storage and is updated in place: a captured CUDA graph re-reads the same addresses
        self.ego.copy_(torch.where(m[:, None], ego, self.ego))
        self.oth.copy_(torch.where(m[:, None, None], oth, self.oth))
        self.t.copy_(torch.where(m, torch.zeros_like(self.t), self.t))
        self.crashed.copy_(self.crashed & ~m)
        return self.observe()

    def observe(self) -> torch.Tensor:
        """[B, V, 8] Kinematics rows (presence, x, y, vx, vy, heading, sin_h, cos_h), ego first, others sorted by
        distance to the ego (config/config.py:13, cfg.yaml:4-6)."""
        e, o = self.ego, self.oth
        dist_ = torch.hypot(o[..., 0] - e[:, None, 0], o[..., 1] - e[:, None, 1])
        idx = torch.argsort(dist_, dim=1)
        o = torch.gather(o, 1, idx[..., None].expand(-1, -1, 4))
        x = torch.cat([e[:, None, 0], o[..., 0]], dim=1)
        y = torch.cat([e[:, None, 1], o[..., 1]], dim=1)
        spd = torch.cat([e[:, None, 3], o[..., 2]], dim=1)
        hd = torch.cat([e[:, None, 2], o[..., 3]], dim=1)
        sh, ch = torch.sin(hd), torch.cos(hd)
        return torch.stack([torch.ones_like(x), x, y, spd * ch, spd * sh, hd, sh, ch], dim=-1).contiguous()

    def step(self, action: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
        """action [B, 2] = (accel, steer) as the caller sends it.  Returns (obs, reward, done, info); finished
        environments are reset in place (VecEnv semantics) and flagged in `done`; info["terminal_observation"]
        holds the observation before that reset (meaningful in the rows where `done`)."""
        if self.fused:
            return self._step_fused(action)
        a = action.to(self.device).float()
        if self.action_scaling == "sb3_raw":               # agents/a2c_mpc.py:151-153 -> ContinuousAction clip + lmap (quirk Q8)
            acc = a[:, 0].clamp(-1, 1) * 5.0
            steer = a[:, 1].clamp(-1, 1) * (math.pi / 4)
        else:
            acc, steer = a[:, 0].clamp(-5, 5), a[:, 1].clamp(-math.pi / 4, math.pi / 4)
        x, y, th, v = self.ego.unbind(1)
        beta = torch.atan(0.5 * torch.tan(steer))
        for _ in range(self.sub):                          # highway-env Vehicle.step at the simulation frequency
            x = x + v * torch.cos(th + beta) * self.dt_sim
            y = y + v * torch.sin(th + beta) * self.dt_sim
            th = th + v * torch.sin(beta) / 2.5 * self.dt_sim
            v = (v + acc * self.dt_sim).clamp(-40.0, 40.0)
        th = torch.atan2(torch.sin(th), torch.cos(th))
        self.ego.copy_(torch.stack([x, y, th, v], dim=1))
        o = self.oth
        step_len = o[..., 2] * (self.sub * self.dt_sim)
        ox = o[..., 0] + step_len * torch.cos(o[..., 3])
        oy = o[..., 1] + step_len * torch.sin(o[..., 3])
        gone = (ox.abs() > 90) | (oy.abs() > 90)           # _clear_vehicles + _spawn_vehicle stand-in: re-enter at the map edge
        far = self._spawn_others()
        fx = torch.where(far[..., 3] == 0.0, -80.0, torch.where(far[..., 3] == math.pi, 80.0, far[..., 0]))
        fy = torch.where(far[..., 3] == -math.pi / 2, 80.0, torch.where(far[..., 3] == math.pi / 2, -80.0, far[..., 1]))
        ox, oy = torch.where(gone, fx, ox), torch.where(gone, fy, oy)
        self.oth.copy_(torch.stack([ox, oy, torch.where(gone, far[..., 2], o[..., 2]), torch.where(gone, far[..., 3], o[..., 3])], dim=-1))
        self.t += 1
        crashed = (torch.hypot(ox - x[:, None], oy - y[:, None]) < 2.5).any(dim=1)
        arrived = (x <= self.ref_xy[-2, 0]) & ((y - self.ref_xy[-1, 1]).abs() < 4.0)
        speed_r = ((v - 7.0) / 2.0).clamp(0, 1)            # lmap(speed, [7, 9], [0, 1]) clipped
        reward = -5.0 * crashed.float() + 1.0 * speed_r
        reward = torch.where(arrived, torch.ones_like(reward), reward)
        truncated = self.t >= self.duration_steps
        done = crashed | arrived | truncated
        self.crashed.copy_(crashed)
        info = {"crashed": crashed, "arrived": arrived, "truncated": truncated & ~(crashed | arrived), "speed": v.clone(),
                "terminal_observation": self.observe()}
        obs = self.reset(done)
        return obs, reward, done, info


def _fused_step(self, action: torch.Tensor):
    """`step` as one launch of k_env_step (csrc/mpc_env.cu) on the current stream; no host synchronisation."""
    C_ = self._capi.C
    B, V, dev = self.B, self.V, self.device
    a = action.to(device=dev, dtype=torch.float32).contiguous()
    obs = torch.empty(B, V, 8, device=dev)
    term = torch.empty(B, V, 8, device=dev)
    reward = torch.empty(B, device=dev)
    speed = torch.empty(B, device=dev)
    flags = torch.empty(4, B, dtype=torch.bool, device=dev)              # done, crashed, arrived, truncated
    args = self._capi.MpcEnvStep(
        ego=self.ego.data_ptr(), others=self.oth.data_ptr(), t=self.t.data_ptr(), crashed_state=self.crashed.data_ptr(),
        counter=self._ctr.data_ptr(), action=a.data_ptr(), obs=obs.data_ptr(), terminal_obs=term.data_ptr(),
        reward=reward.data_ptr(), done=flags[0].data_ptr(), crashed=flags[1].data_ptr(), arrived=flags[2].data_ptr(),
        truncated=flags[3].data_ptr(), speed=speed.data_ptr(), B=B, n_others=self.M, substeps=self.sub,
        duration_steps=int(self.duration_steps), raw_action=int(self.action_scaling == "sb3_raw"), dt_sim=float(self.dt_sim),
        arrive_x=self._arrive[0], arrive_y=self._arrive[1])
    rc = self._lib.mpc_env_step(C_.byref(args), torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        raise RuntimeError(f"mpc_env_step failed with code {rc}")
    self._ctr += _i64(10 * 0xD1B54A32D192ED03)                            # the ten draw calls of one step (spawn 4, reset 2 + 4)
    self._keep = (a, obs, term)
    info = {"crashed": flags[1], "arrived": flags[2], "truncated": flags[3], "speed": speed, "terminal_observation": term}
    return obs, reward, flags[0], info


BatchedIntersectionEnv._step_fused = _fused_step


class _MlpExtractor(nn.Module):
    def __init__(self, obs_dim: int, width: int = 64):
        super().__init__()
        self.policy_net = nn.Sequential(nn.Linear(obs_dim, width), nn.Tanh(), nn.Linear(width, width), nn.Tanh())
        self.value_net = nn.Sequential(nn.Linear(obs_dim, width), nn.Tanh(), nn.Linear(width, width), nn.Tanh())


class ActorCritic(nn.Module):
    """SB3 `MlpPolicy` as the reference builds it (trainers/trainer_utils.py:6-44: ActorCriticPolicy over
    Box(-1, 1, (action_dim,))): separate 64-64 tanh towers, linear action / value heads.  Module and
    parameter names are SB3's, so `state_dict()` is interchangeable with the `policy.pth` inside the
    reference's zip checkpoints (weights/v0/*.zip; see checkpoint.py).

    use_sde=False: state-independent log-std [action_dim] (A2C_MPC, cfg.yaml:43).
    use_sde=True : generalised state-dependent exploration (PPO_MPC default, agents/ppo_mpc.py:114):
                   log_std [64, action_dim]; noise = latent_pi @ theta with theta ~ N(0, exp(log_std)^2) drawn
                   per environment by `reset_noise`; likelihood N(mean, sqrt(latent_pi^2 @ exp(log_std)^2))."""

    def __init__(self, obs_dim: int, action_dim: int = 1, use_sde: bool = False, log_std_init: float = 0.0):
        super().__init__()
        self.use_sde, self.action_dim = bool(use_sde), int(action_dim)
        self.mlp_extractor = _MlpExtractor(obs_dim)
        self.action_net = nn.Linear(64, action_dim)
        self.value_net = nn.Linear(64, 1)
        shape = (64, action_dim) if use_sde else (action_dim,)
        self.log_std = nn.Parameter(torch.full(shape, float(log_std_init)))
        self._theta = None                                    # [B, 64, A] exploration matrices (gSDE)
        self._theta_pinned = False                            # a CUDA graph holds the buffer's address

    def value(self, obs):
        return self.value_net(self.mlp_extractor.value_net(obs)).squeeze(-1)

    def reset_noise(self, n_envs: int) -> None:
        """Draws the exploration matrices of the ROLLOUT (one per environment).  The buffer is only ever updated in
        place once it exists: a captured CUDA graph reads its address, so re-binding it (e.g. for a minibatch of a
        different size) would leave the graph reading freed memory.  `evaluate_actions` never uses the matrices, so the
        update loop has no reason to call this."""
        if self.use_sde:
            std = self.log_std.detach().exp()
            theta = torch.randn(n_envs, *std.shape, device=std.device) * std
            if self._theta is None:
                self._theta = theta
            elif self._theta.shape == theta.shape and self._theta.device == theta.device:
                self._theta.copy_(theta)                      # same storage: a captured CUDA graph keeps reading it
            elif self._theta_pinned:
                raise RuntimeError("reset_noise with a different shape after a CUDA graph captured the noise buffer")
            else:
                self._theta = theta

    def dist(self, obs):
        latent = self.mlp_extractor.policy_net(obs)
        mean = self.action_net(latent)
        if self.use_sde:
            var = (latent.detach() ** 2) @ (self.log_std.exp() ** 2)        # SB3: no gradient through the features
            return torch.distributions.Normal(mean, torch.sqrt(var + 1e-6), validate_args=False), latent
        return torch.distributions.Normal(mean, self.log_std.exp().expand_as(mean), validate_args=False), latent   # validation would sync

    def forward(self, obs, deterministic: bool = False):
        d, latent = self.dist(obs)
        if deterministic:
            a = d.mean
        elif self.use_sde:
            if self._theta is None or self._theta.shape[0] != obs.shape[0]:
                self.reset_noise(obs.shape[0])
            a = d.mean + torch.bmm(latent.detach().unsqueeze(1), self._theta).squeeze(1)
        else:
            a = d.mean + d.stddev * torch.randn_like(d.mean)   # not d.sample(): torch.normal(tensor std) checks std >= 0 on the host
        return a, self.value(obs), d.log_prob(a).sum(-1)

    def evaluate_actions(self, obs, actions):
        d, _ = self.dist(obs)
        return self.value(obs), d.log_prob(actions).sum(-1), d.entropy().sum(-1)


class _MPCRollout:
    """Rollout shared by the batched A2C-MPC and PPO-MPC: policy(obs) -> RL action -> MPC (reference speed in
    version "v0", the three objective weights in "v1") -> env.step(raw (a, delta)); the buffer stores the RL
    action (agents/a2c_mpc.py:111-180, agents/ppo_mpc.py:353-483).  Time-limit truncations bootstrap with the
    value of the terminal observation as SB3 does (agents/ppo_mpc.py:451-461).

    graph=True (CUDA only): the whole transition -- policy forward, the MPC's two kernels, the env step, the
    buffer writes -- is captured once in a CUDA graph and replayed n_steps times per rollout; graph=False runs
    it eagerly with per-phase wall-clock shares in `stats`."""

    clip_rl_action = False      # PPO_MPC clips the sample to the Box before the MPC sees it (agents/ppo_mpc.py:399-408);
                                # A2C_MPC hands the raw Gaussian sample over

    def _setup(self, env, mpc, version, action_dim, n_steps, gamma, gae_lambda, use_sde, seed, graph=False):
        if version not in ("v0", "v1"):
            raise ValueError("version must be 'v0' (RL sets the reference speed) or 'v1' (RL sets the MPC weights)")
        self.env, self.mpc, self.version = env, mpc, version
        self.n_steps, self.gamma, self.lam = n_steps, gamma, gae_lambda
        self.action_dim = action_dim if action_dim is not None else (1 if version == "v0" else 3)
        torch.manual_seed(seed)
        self.policy = ActorCritic(env.V * 8, self.action_dim, use_sde=use_sde).to(env.device)
        if dist.is_available() and dist.is_initialized():
            for p in self.policy.parameters():
                dist.broadcast(p.data, 0)
        B, T, dev = env.B, n_steps, env.device
        self.obs = env.observe().clone()                     # static storage: the captured transition updates it in place
        self.episode_start = torch.ones(B, dtype=torch.bool, device=dev)
        self._buf = {k: torch.zeros(T, B, device=dev) for k in ("rew", "val", "logp", "done")}
        self._buf["act"] = torch.zeros(T, B, self.action_dim, device=dev)
        self._obs_buf = torch.zeros(T, B, env.V * 8, device=dev)
        self._row = torch.zeros(1, dtype=torch.int64, device=dev)      # row of the rollout buffer the next transition fills
        self.graph = bool(graph) and dev.type == "cuda"
        self._cuda_graph = None
        self.num_timesteps = 0
        self.stats = {"mpc_s": 0.0, "env_s": 0.0, "policy_s": 0.0, "update_s": 0.0, "rollout_s": 0.0, "steps": 0}

    def mpc_action(self, obs, rl_action, reset_mask=None):
        a = rl_action.clamp(-1.0, 1.0) if self.clip_rl_action else rl_action
        if self.version == "v0":
            return self.mpc.predict_batch(obs, ref_speed=a[:, :1].contiguous(), reset_mask=reset_mask)
        return self.mpc.predict_batch(obs, weights=a[:, :3].contiguous(), reset_mask=reset_mask)

    @torch.no_grad()
    def _transition(self, timed: bool = False):
        """One policy -> MPC -> env transition written to row `_row` of the rollout buffers.  No host
        synchronisation unless `timed` (per-phase wall clock for the eager profile)."""
        B, dev = self.env.B, self.env.device
        sync = (lambda: torch.cuda.synchronize(dev)) if (timed and dev.type == "cuda") else (lambda: None)
        t0 = time.perf_counter()
        flat = self.obs.reshape(B, -1)
        a, v, lp = self.policy(flat)
        sync()
        t1 = time.perf_counter()
        mpc_action = self.mpc_action(self.obs, a, self.episode_start)   # the latch of finished envs is cleared
        sync()
        t2 = time.perf_counter()
        new_obs, rew, done, info = self.env.step(mpc_action)
        tv = self.policy.value(info["terminal_observation"].reshape(B, -1))
        rew = rew + self.gamma * tv * info["truncated"].float()         # time-limit bootstrap (agents/ppo_mpc.py:451-461)
        sync()
        t3 = time.perf_counter()
        r = self._row
        self._obs_buf.index_copy_(0, r, flat.unsqueeze(0))
        self._buf["act"].index_copy_(0, r, a.unsqueeze(0))
        self._buf["rew"].index_copy_(0, r, rew.unsqueeze(0))
        self._buf["val"].index_copy_(0, r, v.unsqueeze(0))
        self._buf["logp"].index_copy_(0, r, lp.unsqueeze(0))
        self._buf["done"].index_copy_(0, r, done.float().unsqueeze(0))
        self.obs.copy_(new_obs)
        self.episode_start.copy_(done)
        self._row += 1
        if timed:
            self.stats["policy_s"] += t1 - t0
            self.stats["mpc_s"] += t2 - t1
            self.stats["env_s"] += t3 - t2

    def _capture(self):
        """Captures `_transition` into a CUDA graph (a few hundred small launches + the two MPC kernels become
        one replay): warm up on a side stream, then capture.  The MPC launches are recorded through the C ABI
        on torch's capturing stream; all operands live in static storage."""
        dev = self.env.device
        self.policy.reset_noise(self.env.B)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._row.zero_()
                self._transition()
            self._row.zero_()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._transition()
        self._cuda_graph = g
        self.policy._theta_pinned = True                    # the graph reads the noise buffer by address from now on

    def collect_rollouts(self):
        B, T, dev = self.env.B, self.n_steps, self.env.device
        if self.graph and self._cuda_graph is None:
            self._capture()
        self.policy.reset_noise(B)                          # sde_sample_freq = -1: once per rollout
        self._row.zero_()
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(T):
            if self._cuda_graph is not None:
                self._cuda_graph.replay()
            else:
                self._transition(timed=True)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        self.stats["rollout_s"] += time.perf_counter() - t0
        self.stats["steps"] += B * T
        self.num_timesteps += B * T
        buf, obs_buf = self._buf, self._obs_buf
        with torch.no_grad():
            last_v = self.policy.value(self.obs.reshape(B, -1))
        adv = torch.zeros(T, B, device=dev)
        gae = torch.zeros(B, device=dev)
        for t in reversed(range(T)):                        # SB3 RolloutBuffer.compute_returns_and_advantage
            nv = last_v if t == T - 1 else buf["val"][t + 1]
            nonterminal = 1.0 - buf["done"][t]
            delta = buf["rew"][t] + self.gamma * nv * nonterminal - buf["val"][t]
            gae = delta + self.gamma * self.lam * nonterminal * gae
            adv[t] = gae
        return obs_buf, buf, adv, adv + buf["val"]

    def _step_optimizer(self, loss):
        self.opt.zero_grad()
        loss.backward()
        if dist.is_available() and dist.is_initialized():   # data-parallel update over the env shards
            w = dist.get_world_size()
            for p in self.policy.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad)
                    p.grad /= w
        nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
        self.opt.step()

    @torch.no_grad()
    def predict(self, obs, deterministic: bool = True):
        """Inference as `BaseTrainer.predict` does it (trainers/trainer.py:345-373): policy -> MPC -> (a, delta)."""
        a, _, _ = self.policy(obs.reshape(obs.shape[0], -1), deterministic=deterministic)
        return self.mpc_action(obs, a)


class A2CMPC(_MPCRollout):
    """Batched counterpart of `A2C_MPC` (agents/a2c_mpc.py) with the hyper-parameters of cfg.yaml:30-45."""

    def __init__(self, env: BatchedIntersectionEnv, mpc, n_steps: int = 64, lr: float = 7e-4, gamma: float = 0.99,
                 gae_lambda: float = 1.0, ent_coef: float = 0.0, vf_coef: float = 0.5, max_grad_norm: float = 0.5,
                 rms_prop_eps: float = 1e-5, seed: int = 0, version: str = "v0", action_dim: Optional[int] = None,
                 use_sde: bool = False, graph: bool = False):
        self._setup(env, mpc, version, action_dim, n_steps, gamma, gae_lambda, use_sde, seed, graph)
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        self.opt = torch.optim.RMSprop(self.policy.parameters(), lr=lr, alpha=0.99, eps=rms_prop_eps)

    def train_step(self) -> Dict[str, float]:
        obs_buf, buf, adv, ret = self.collect_rollouts()
        t0 = time.perf_counter()
        flat = obs_buf.reshape(-1, obs_buf.shape[-1])
        values, logp, entropy = self.policy.evaluate_actions(flat, buf["act"].reshape(-1, self.action_dim))
        policy_loss = -(adv.reshape(-1) * logp).mean()      # agents/a2c_mpc.py:182-226 (normalize_advantage false)
        value_loss = torch.nn.functional.mse_loss(ret.reshape(-1), values)
        entropy_loss = -entropy.mean()
        loss = policy_loss + self.ent_coef * entropy_loss + self.vf_coef * value_loss
        self._step_optimizer(loss)
        if self.env.device.type == "cuda":
            torch.cuda.synchronize(self.env.device)
        self.stats["update_s"] += time.perf_counter() - t0
        return {"loss": float(loss.detach()), "policy_loss": float(policy_loss.detach()), "value_loss": float(value_loss.detach()),
                "mean_reward": float(buf["rew"].mean()), "done_rate": float(buf["done"].mean())}


class PPOMPC(_MPCRollout):
    """Batched counterpart of `PPO_MPC` (agents/ppo_mpc.py): clipped-surrogate PPO around the MPC, gSDE on by
    default (agents/ppo_mpc.py:114), hyper-parameters of cfg.yaml:62-87 (lr 3e-4, 10 epochs, clip 0.2,
    GAE 0.95, advantages normalised per minibatch, vf_coef 0.5, max_grad_norm 0.5, Adam).  `batch_size` is the
    minibatch size of the update (agents/ppo_mpc.py:213-333); with B environments a rollout holds
    B * n_steps samples, so the default scales the reference's 256 by the number of environments."""

    clip_rl_action = True

    def __init__(self, env: BatchedIntersectionEnv, mpc, n_steps: int = 256, batch_size: Optional[int] = None,
                 n_epochs: int = 10, lr: float = 3e-4, gamma: float = 0.99, gae_lambda: float = 0.95,
                 clip_range: float = 0.2, clip_range_vf: Optional[float] = None, normalize_advantage: bool = True,
                 ent_coef: float = 0.0, vf_coef: float = 0.5, max_grad_norm: float = 0.5, target_kl: Optional[float] = None,
                 use_sde: bool = True, seed: int = 0, version: str = "v0", action_dim: Optional[int] = None,
                 graph: bool = False):
        self._setup(env, mpc, version, action_dim, n_steps, gamma, gae_lambda, use_sde, seed, graph)
        self.batch_size = batch_size if batch_size is not None else 256 * env.B
        self.n_epochs, self.clip_range, self.clip_range_vf = n_epochs, clip_range, clip_range_vf
        self.normalize_advantage, self.target_kl = normalize_advantage, target_kl
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=lr, eps=1e-5)
        self.gen = torch.Generator(device=env.device)
        self.gen.manual_seed(seed + 1)

    def train_step(self) -> Dict[str, float]:
        obs_buf, buf, adv, ret = self.collect_rollouts()
        t0 = time.perf_counter()
        n = obs_buf.shape[0] * obs_buf.shape[1]
        obs_f, act_f = obs_buf.reshape(n, -1), buf["act"].reshape(n, self.action_dim)
        adv_f, ret_f, val_f, logp_f = adv.reshape(n), ret.reshape(n), buf["val"].reshape(n), buf["logp"].reshape(n)
        bs = min(self.batch_size, n)
        last = {}
        stop = False
        for _ in range(self.n_epochs):
            perm = torch.randperm(n, device=obs_f.device, generator=self.gen)
            for lo in range(0, n, bs):
                idx = perm[lo:lo + bs]
                values, logp, entropy = self.policy.evaluate_actions(obs_f[idx], act_f[idx])
                a = adv_f[idx]
                if self.normalize_advantage and a.numel() > 1:
                    a = (a - a.mean()) / (a.std() + 1e-8)
                ratio = torch.exp(logp - logp_f[idx])
                policy_loss = -torch.min(a * ratio, a * ratio.clamp(1 - self.clip_range, 1 + self.clip_range)).mean()
                vp = values if self.clip_range_vf is None else val_f[idx] + (values - val_f[idx]).clamp(-self.clip_range_vf, self.clip_range_vf)
                value_loss = torch.nn.functional.mse_loss(ret_f[idx], vp)
                entropy_loss = -entropy.mean()
                loss = policy_loss + self.ent_coef * entropy_loss + self.vf_coef * value_loss
                with torch.no_grad():
                    log_ratio = logp - logp_f[idx]
                    kl_t = ((log_ratio.exp() - 1) - log_ratio).mean()
                    if self.target_kl is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                        dist.all_reduce(kl_t, op=dist.ReduceOp.AVG)      # every rank must take the same branch below
                    approx_kl = float(kl_t)
                    clip_frac = float(((ratio - 1).abs() > self.clip_range).float().mean())
                if self.target_kl is not None and approx_kl > 1.5 * self.target_kl:
                    stop = True
                    break
                self._step_optimizer(loss)
                last = {"loss": float(loss.detach()), "policy_loss": float(policy_loss.detach()),
                        "value_loss": float(value_loss.detach()), "approx_kl": approx_kl, "clip_fraction": clip_frac}
            if stop:
                break
        if self.env.device.type == "cuda":
            torch.cuda.synchronize(self.env.device)
        self.stats["update_s"] += time.perf_counter() - t0
        last.update({"mean_reward": float(buf["rew"].mean()), "done_rate": float(buf["done"].mean())})
        return last


# ---------------------------------------------------------------------------------------------------------------
# Reference-signature wrappers: `A2C_MPC` / `PPO_MPC` with the constructor keywords and the `learn` /
# `collect_rollouts` / `train` / `predict` method names of agents/a2c_mpc.py:53-244 and agents/ppo_mpc.py:94-483,
# accepting n_envs > 1 (the reference is hard-wired to `self.env.envs[0]`, a2c_mpc.py:104 / ppo_mpc.py:198).
# stable-baselines3 is not installable here, so these are NOT SB3 subclasses: `policy` must be "MlpPolicy" (the
# 64-64 tanh actor-critic the reference builds, trainers/trainer_utils.py:6-44), `env` a BatchedIntersectionEnv, and
# SB3-only arguments (tensorboard_log, policy_kwargs, rollout_buffer_class, callbacks ...) are accepted and ignored.
# ---------------------------------------------------------------------------------------------------------------
def _make_batched_mpc(env, pure_mpc_cfg, collision_check=True):
    from .agent import BatchedPureMPC
    return BatchedPureMPC(pure_mpc_cfg, vehicles_count=env.V, max_batch=env.B, device=env.device,
                          collision_check=collision_check)


class _SB3Surface:
    def learn(self, total_timesteps: int, callback=None, log_interval: int = 100, tb_log_name: str = "", reset_num_timesteps: bool = True,
              progress_bar: bool = False):
        """OnPolicyAlgorithm.learn: alternate collect_rollouts / train until `total_timesteps` env steps were taken."""
        start = 0 if reset_num_timesteps else self.num_timesteps
        if reset_num_timesteps:
            self.num_timesteps = 0
        while self.num_timesteps - start < total_timesteps:
            self.last_log = self.train_step()
        return self

    def train(self):
        """SB3 splits collect_rollouts() and train(); the batched loop fuses them in train_step()."""
        self.last_log = self.train_step()


class A2C_MPC(_SB3Surface, A2CMPC):
    """agents/a2c_mpc.py:53-109 keyword surface over the batched A2C-MPC loop."""

    def __init__(self, mpcrl_cfg: Dict, version: str, pure_mpc_cfg: Dict, policy="MlpPolicy", env=None, learning_rate: float = 7e-4,
                 n_steps: int = 64, gamma: float = 0.99, gae_lambda: float = 1.0, ent_coef: float = 0.0, vf_coef: float = 0.5,
                 max_grad_norm: float = 0.5, rms_prop_eps: float = 1e-5, use_rms_prop: bool = True, use_sde: bool = False,
                 sde_sample_freq: int = -1, normalize_advantage: bool = False, tensorboard_log=None, policy_kwargs=None,
                 verbose: int = 0, seed: Optional[int] = None, device="auto", _init_setup_model: bool = True, graph: bool = False):
        if policy not in ("MlpPolicy", None) or not use_rms_prop or sde_sample_freq != -1 or normalize_advantage:
            raise NotImplementedError("only the reference's own settings: MlpPolicy, RMSprop, sde_sample_freq -1, normalize_advantage False")
        self.mpcrl_cfg, self.normalize_advantage = mpcrl_cfg, normalize_advantage
        self.mpc_agent = _make_batched_mpc(env, pure_mpc_cfg)
        A2CMPC.__init__(self, env, self.mpc_agent, n_steps=n_steps, lr=learning_rate, gamma=gamma, gae_lambda=gae_lambda,
                        ent_coef=ent_coef, vf_coef=vf_coef, max_grad_norm=max_grad_norm, rms_prop_eps=rms_prop_eps,
                        seed=0 if seed is None else seed, version=version,
                        action_dim=(mpcrl_cfg or {}).get("action_space_dim"), use_sde=use_sde, graph=graph)


class PPO_MPC(_SB3Surface, PPOMPC):
    """agents/ppo_mpc.py:94-200 keyword surface over the batched PPO-MPC loop (`use_collision_avoidance=False` selects
    the pure_mpc_no_collision flow, ppo_mpc.py:186-199)."""

    def __init__(self, mpcrl_cfg: Dict, version: str, pure_mpc_cfg: Dict, policy="MlpPolicy", env=None, use_collision_avoidance: bool = True,
                 learning_rate: float = 3e-4, n_steps: int = 2048, batch_size: int = 64, n_epochs: int = 10, gamma: float = 0.99,
                 gae_lambda: float = 0.95, clip_range: float = 0.2, clip_range_vf=None, normalize_advantage: bool = True,
                 ent_coef: float = 0.0, vf_coef: float = 0.5, max_grad_norm: float = 0.5, use_sde: bool = True, sde_sample_freq: int = -1,
                 rollout_buffer_class=None, rollout_buffer_kwargs=None, target_kl: Optional[float] = None, stats_window_size: int = 100,
                 tensorboard_log=None, policy_kwargs=None, verbose: int = 0, seed: Optional[int] = None, device="auto",
                 _init_setup_model: bool = True, graph: bool = False):
        if policy not in ("MlpPolicy", None) or sde_sample_freq != -1:
            raise NotImplementedError("only the reference's own settings: MlpPolicy, sde_sample_freq -1")
        self.mpcrl_cfg = mpcrl_cfg
        self.mpc_agent = _make_batched_mpc(env, pure_mpc_cfg, collision_check=use_collision_avoidance)
        PPOMPC.__init__(self, env, self.mpc_agent, n_steps=n_steps, batch_size=batch_size, n_epochs=n_epochs, lr=learning_rate,
                        gamma=gamma, gae_lambda=gae_lambda, clip_range=clip_range, clip_range_vf=clip_range_vf,
                        normalize_advantage=normalize_advantage, ent_coef=ent_coef, vf_coef=vf_coef, max_grad_norm=max_grad_norm,
                        target_kl=target_kl, use_sde=use_sde, seed=0 if seed is None else seed, version=version,
                        action_dim=(mpcrl_cfg or {}).get("action_space_dim"), graph=graph)
