"""the reference's model comparison (main/model_comparison.py:40-175) over many seeded
episodes at once.  One episode = reset -> step until arrival / crash / time limit (150 policy steps in the
reference script); per model it reports the same five aggregates: success rate, collision rate, mean steps,
mean speed, mean travel time.

Controllers are callables `(obs [B,V,8], episode_start [B] bool) -> action [B,2]` returning what the
reference script hands to `env.step`:
  * "pure_mpc" / "pure_mpc_no_collision": the MPC's (a, delta) as is (main/model_comparison.py:50-53);
  * "mpcrl": policy -> MPC -> (a / 5, delta / (pi/3)) (main/model_comparison.py:54-56).
The environment is whatever implements `observe / step / reset(mask)` like rl.BatchedIntersectionEnv (raw
SB3 action path: clip to [-1, 1], scale to +-5 m/s^2 and +-pi/4).
"""
from __future__ import annotations

import math
from typing import Callable, Dict

import torch


def pure_mpc_controller(mpc) -> Callable:
    return lambda obs, start: mpc.predict_batch(obs, reset_mask=start)


def mpcrl_controller(algo, deterministic: bool = False) -> Callable:
    """`algo`: rl.A2CMPC / rl.PPOMPC (or anything with .policy and .mpc_action).  The reference evaluates with
    `model.predict(observation, False)`, i.e. stochastic actions; pass deterministic=True for the mean."""
    scale = torch.tensor([1.0 / 5.0, 1.0 / (math.pi / 3)])

    @torch.no_grad()
    def ctrl(obs, start):
        a, _, _ = algo.policy(obs.reshape(obs.shape[0], -1), deterministic=deterministic)
        return algo.mpc_action(obs, a, start) * scale.to(obs.device)
    return ctrl


@torch.no_grad()
def evaluate(controller: Callable, env, n_episodes: int, max_steps: int = 150) -> Dict[str, float]:
    """Runs ceil(n_episodes / B) complete episodes in every environment (all in parallel, each restarting in place) and
    aggregates them; `episodes` in the result is the number counted (a multiple of B, >= n_episodes)."""
    B, dev = env.B, env.device
    obs = env.reset()
    start = torch.ones(B, dtype=torch.bool, device=dev)
    steps = torch.zeros(B, device=dev)
    speed_sum = torch.zeros(B, device=dev)
    tot = {"episodes": 0, "successes": 0, "collisions": 0, "total_steps": 0.0, "total_speed": 0.0}
    dt = env.sub * env.dt_sim
    guard = 0
    # every environment contributes the same number of complete episodes (its first `quota`): counting the first
    # n_episodes episodes to FINISH anywhere would over-sample short ones (early crashes) -- the reference runs n
    # complete episodes one after another (main/model_comparison.py:40-105)
    quota = -(-int(n_episodes) // B)
    done_count = torch.zeros(B, dtype=torch.long, device=dev)
    while int(done_count.min()) < quota:
        action = controller(obs, start)
        speed_sum += env.ego[:, 3]                          # speed before the step (main/model_comparison.py:62-63)
        obs, _, done, info = env.step(action)
        steps += 1
        over = (steps >= max_steps) & ~done                # the script's own step limit (no_steps), env still running
        if bool(over.any()):
            obs = env.reset(over)
        fin = done | over
        if bool(fin.any()):
            cnt = fin & (done_count < quota)
            idx = torch.nonzero(cnt).flatten()
            tot["episodes"] += int(idx.numel())
            tot["successes"] += int(info["arrived"][idx].sum())
            tot["collisions"] += int(info["crashed"][idx].sum())
            tot["total_steps"] += float(steps[idx].sum())
            tot["total_speed"] += float((speed_sum[idx] / steps[idx]).sum())
            done_count += fin.long()
            steps[fin] = 0
            speed_sum[fin] = 0
        start = fin
        guard += 1
        if guard > 100 * max_steps * quota:
            raise RuntimeError("evaluation did not finish: the environment never ends an episode")
    n = max(tot["episodes"], 1)
    return {"episodes": tot["episodes"], "success_rate": tot["successes"] / n, "collision_rate": tot["collisions"] / n,
            "avg_steps": tot["total_steps"] / n, "avg_speed": tot["total_speed"] / n, "avg_time": tot["total_steps"] / n * dt}


def compare(models: Dict[str, Callable], make_env: Callable, n_episodes: int, max_steps: int = 150) -> Dict[str, Dict[str, float]]:
    """`models`: name -> controller; `make_env()` builds a fresh identically seeded environment per model, as the
    reference re-uses one seeded env config for every model (main/model_comparison.py:108-160)."""
    return {name: evaluate(ctrl, make_env(), n_episodes, max_steps) for name, ctrl in models.items()}
