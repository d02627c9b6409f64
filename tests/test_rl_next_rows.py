"""batched synthetic env + batched A2C-MPC loop.  CPU: env rules and the rollout /
update plumbing with a stub controller; GPU: the real MPC in the loop."""
import math

import numpy as np
import pytest
import torch


class _StubMPC:
    """Constant controller with the BatchedPureMPC.predict_batch signature (CPU plumbing test only)."""

    def predict_batch(self, obs, ref_speed=None, weights=None, reset_mask=None):
        a = torch.zeros(obs.shape[0], 2)
        a[:, 0] = 0.2
        return a


def test_env_rules_cpu():
    from mpc_rl_for_avs_b200.rl import BatchedIntersectionEnv
    env = BatchedIntersectionEnv(16, 9, device="cpu", seed=3)
    obs = env.observe()
    assert obs.shape == (16, 10, 8) and (obs[:, :, 0] == 1).all()
    d = torch.hypot(obs[:, 1:, 1] - obs[:, :1, 1], obs[:, 1:, 2] - obs[:, :1, 2])
    assert (d[:, 1:] >= d[:, :-1]).all()                                   # others sorted by distance (cfg.yaml:4-6)
    assert torch.allclose(obs[:, 0, 5], torch.full((16,), -math.pi / 2))
    # raw SB3 action path (quirk Q8): accel clipped to +-1 then x5
    v0 = env.ego[:, 3].clone()
    obs2, rew, done, info = env.step(torch.tensor([[7.0, 0.0]]).repeat(16, 1))
    assert torch.allclose(env.ego[:, 3][~done], (v0 + 0.5)[~done], atol=1e-5)
    assert ((rew <= 1.0) & (rew >= -5.0)).all()
    # a vehicle parked on the ego's nose crashes it: reward -5, episode ends and the env is reset in place
    env.oth[0, 0, :2] = env.ego[0, :2] + torch.tensor([0.0, -1.0])
    env.oth[0, 0, 2] = 0.0
    _, rew, done, info = env.step(torch.zeros(16, 2))
    assert bool(info["crashed"][0]) and bool(done[0]) and rew[0] <= -4.0 and int(env.t[0]) == 0
    # time limit
    env2 = BatchedIntersectionEnv(4, 9, device="cpu", seed=4, duration_steps=3)
    for _ in range(3):
        _, _, done, info = env2.step(torch.tensor([[-1.0, 0.0]]).repeat(4, 1))
    assert bool((info["truncated"] | info["crashed"]).all())
    # seeded
    a = BatchedIntersectionEnv(8, 9, device="cpu", seed=7).observe()
    b = BatchedIntersectionEnv(8, 9, device="cpu", seed=7).observe()
    assert torch.equal(a, b)


def test_a2c_plumbing_cpu():
    from mpc_rl_for_avs_b200.rl import A2CMPC, BatchedIntersectionEnv
    env = BatchedIntersectionEnv(8, 9, device="cpu", seed=1)
    algo = A2CMPC(env, _StubMPC(), n_steps=5)
    p0 = [p.detach().clone() for p in algo.policy.parameters()]
    log = algo.train_step()
    assert all(np.isfinite(v) for v in log.values())
    assert any(not torch.equal(a, b) for a, b in zip(p0, algo.policy.parameters()))
    assert algo.stats["steps"] == 40


@pytest.mark.gpu
def test_a2c_mpc_in_the_loop_gpu():
    import mpc_rl_for_avs_b200 as pkg
    from mpc_rl_for_avs_b200.rl import A2CMPC, BatchedIntersectionEnv
    B = 256
    env = BatchedIntersectionEnv(B, 9, device="cuda", seed=5)
    mpc = pkg.BatchedPureMPC({"horizon": 16, "weight_speed": 1, "weight_control": 1, "weight_input_diff": 1},
                             vehicles_count=10, max_batch=B, collision_check=True)
    algo = A2CMPC(env, mpc, n_steps=8)
    for _ in range(2):
        log = algo.train_step()
    assert all(np.isfinite(v) for v in log.values())
    assert algo.stats["steps"] == 2 * 8 * B and algo.stats["mpc_s"] > 0
    # the latch of finished environments was cleared through reset_mask
    assert int(mpc.collision_memory[:B].max()) <= 10
