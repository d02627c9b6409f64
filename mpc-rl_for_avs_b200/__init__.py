"""mpc-rl_for_avs_b200 -- B200-native batched nonlinear MPC for the PureMPC_Agent hot path of
SaeedRahmani/MPC-RL_for_AVs.  Import name: `mpc_rl_for_avs_b200` (see the shim at the repo root;
the directory keeps the project's hyphenated name).

Public surface
  BatchedPureMPC            thousands of problems per call, torch CUDA tensors in/out
  PureMPC_Agent             drop-in for agents.pure_mpc.PureMPC_Agent (B = 1, numpy in/out)
  PureMPC_NoCollision_Agent drop-in for agents.pure_mpc_no_collision.PureMPC_Agent
  MPC_Action                agents/utils.py:4-12
  make_scenarios            seeded synthetic intersection observations
  sharding                  env-index sharding + all-gather of actions across ranks
  rl                        (next rows N1/N2) batched synthetic intersection env + batched A2C-MPC / PPO-MPC loops
  evaluation, checkpoint    (next row N4) batched model comparison; SB3-zip policy checkpoints
"""
from .agent import MPC_Action, BatchedPureMPC, PureMPC_Agent, PureMPC_NoCollision_Agent  # noqa: F401
from .scenarios import make_scenarios, reference_path  # noqa: F401
from . import sharding  # noqa: F401
from . import rl  # noqa: F401
from . import evaluation, checkpoint  # noqa: F401
from . import _capi  # noqa: F401

__all__ = ["MPC_Action", "BatchedPureMPC", "PureMPC_Agent", "PureMPC_NoCollision_Agent", "make_scenarios",
           "reference_path", "sharding"]
