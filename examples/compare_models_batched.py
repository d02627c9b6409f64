"""The reference's model comparison (main/model_comparison.py) over many seeded episodes at once:
pure MPC, pure MPC without the collision logic, and an RL-set-reference-speed MPC (policy from an SB3-layout
zip, e.g. the reference's weights/v0/test_a2c_v0.zip, or untrained).

    python examples/compare_models_batched.py [--envs 1024] [--episodes 1024] [--policy path.zip]

Environment: the synthetic stand-in of mpc_rl_for_avs_b200.rl (highway-env is not installable here), so the numbers
are not comparable with the reference's highway-env runs; the metrics and their definitions are the same.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_rl_for_avs_b200 as pkg
from mpc_rl_for_avs_b200 import checkpoint, evaluation
from mpc_rl_for_avs_b200.rl import A2CMPC, BatchedIntersectionEnv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--episodes", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=150)          # no_steps of main/model_comparison.py
    ap.add_argument("--policy", default="")
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    V = 10
    cfg = {"horizon": 16, "weight_speed": 1, "weight_control": 1, "weight_input_diff": 1}
    make_env = lambda: BatchedIntersectionEnv(args.envs, V - 1, device="cuda", seed=args.seed, duration_steps=args.steps)  # noqa: E731
    mpc = pkg.BatchedPureMPC(cfg, vehicles_count=V, max_batch=args.envs, collision_check=True)
    mpc_nc = pkg.BatchedPureMPC(cfg, vehicles_count=V, max_batch=args.envs, collision_check=False)
    mpc_rl = pkg.BatchedPureMPC(cfg, vehicles_count=V, max_batch=args.envs, collision_check=True)
    algo = A2CMPC(make_env(), mpc_rl, n_steps=1)
    if args.policy:
        algo.policy, _ = checkpoint.load_sb3_policy(args.policy, device="cuda")     # gSDE or plain, as saved
    res = evaluation.compare({"pure_mpc": evaluation.pure_mpc_controller(mpc),
                              "pure_mpc_no_collision": evaluation.pure_mpc_controller(mpc_nc),
                              "mpcrl": evaluation.mpcrl_controller(algo)}, make_env, args.episodes, args.steps)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
