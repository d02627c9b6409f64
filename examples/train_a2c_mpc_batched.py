"""BASELINE config 4: A2C-MPC (v0: RL sets the MPC's reference speed) with 1024 vectorised envs, MPC on the GPU.

    python examples/train_a2c_mpc_batched.py [--envs 1024] [--updates 5] [--n-steps 64] [--algo a2c|ppo] [--save out.zip]
    torchrun --nproc-per-node N examples/train_a2c_mpc_batched.py      # env-sharded, gradient all-reduce

Prints one JSON line: env-steps/s and the share of the step time spent in the policy, the MPC, the env and
the A2C update.  The environment is the synthetic stand-in of mpc_rl_for_avs_b200.rl (highway-env is not installable
here); the rollout structure and hyper-parameters are the reference's (agents/a2c_mpc.py, config/cfg.yaml:30-45).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import mpc_rl_for_avs_b200 as pkg
from mpc_rl_for_avs_b200 import checkpoint
from mpc_rl_for_avs_b200.rl import A2CMPC, PPOMPC, BatchedIntersectionEnv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--updates", type=int, default=5)
    ap.add_argument("--n-steps", type=int, default=64)
    ap.add_argument("--horizon", type=int, default=16)         # config/cfg.yaml:90
    ap.add_argument("--warm-start", action="store_true")
    ap.add_argument("--n-starts", type=int, default=1, help="MPC start portfolio (1 = the reference's single cold start; 0 = library default, 4)")
    ap.add_argument("--algo", choices=("a2c", "ppo"), default="a2c")
    ap.add_argument("--save", default="", help="write the policy as an SB3-layout zip")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph: per-phase shares are measured instead")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_others = 9                                               # vehicles_count 10, config/cfg.yaml:2
    env = BatchedIntersectionEnv(args.envs, n_others, device=f"cuda:{local}", seed=1234 + rank)
    cfg = {"horizon": args.horizon, "weight_speed": 1, "weight_control": 1, "weight_input_diff": 1}
    mpc = pkg.BatchedPureMPC(cfg, vehicles_count=n_others + 1, max_batch=args.envs, device=local, collision_check=True, n_starts=args.n_starts)
    Algo = A2CMPC if args.algo == "a2c" else PPOMPC
    algo = Algo(env, mpc, n_steps=args.n_steps, graph=not args.eager)
    algo.train_step()                                          # warm-up (allocator, first launches)
    for k in algo.stats:
        algo.stats[k] = 0
    t0 = time.perf_counter()
    logs = [algo.train_step() for _ in range(args.updates)]
    wall = time.perf_counter() - t0
    s = algo.stats
    if rank == 0:
        print(json.dumps({"metric": f"{args.algo}_mpc_env_steps_per_sec", "value": world * s["steps"] / wall, "n_gpus": world,
                          "envs_per_gpu": args.envs, "n_steps": args.n_steps, "updates": args.updates, "horizon": args.horizon,
                          "mode": "eager (per-phase sync)" if args.eager else "CUDA graph per transition",
                          "share": {k: s[k] / wall for k in (("policy_s", "mpc_s", "env_s", "update_s") if args.eager else ("rollout_s", "update_s"))},
                          "mpc_solves_per_sec_in_loop": (s["steps"] / s["mpc_s"]) if args.eager else None, "last": logs[-1],
                          "data": "synthetic stand-in env (mpc_rl_for_avs_b200.rl)"}))
    if rank == 0 and args.save:
        checkpoint.save_sb3_policy(args.save, algo.policy, {"n_steps": args.n_steps, "num_timesteps": algo.num_timesteps,
                                                           "n_envs": args.envs}, algo.opt)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
