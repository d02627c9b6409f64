"""policy checkpoints in the Stable-Baselines3 zip layout the reference saves
(`model.save(...)` in trainers/trainer.py:326-343,579-591; loaded by agents/a2c_mpc.py:247-335 and
agents/ppo_mpc.py:486-629; examples under weights/v0/*.zip).

An SB3 zip holds `policy.pth` (the ActorCriticPolicy state dict), `policy.optimizer.pth`,
`pytorch_variables.pth`, `data` (JSON; class objects are cloudpickled and base64-encoded next to a
readable repr) and two text files.  SB3 itself is not installed here, so:

  * `load_sb3_policy` reads `policy.pth` + the JSON scalars and fills an `rl.ActorCritic` (same parameter
    names and shapes as SB3's MlpPolicy with two 64-unit tanh layers per tower) — a reference checkpoint
    drives the batched loop / the evaluation harness unchanged;
  * `save_sb3_policy` writes the same member files with the tensors and the plain hyper-parameters.  The
    cloudpickled class entries SB3's own `load` wants (policy_class, spaces, schedules) cannot be produced
    without SB3: loading our zip back into SB3 needs `custom_objects` for those keys, or
    `policy.load_state_dict(torch.load(policy.pth))` on a policy the reference constructs.
"""
from __future__ import annotations

import io
import json
import zipfile
from typing import Dict, Optional, Tuple

import torch

from .rl import ActorCritic

_PLAIN_KEYS = ("use_sde", "sde_sample_freq", "n_steps", "gamma", "gae_lambda", "ent_coef", "vf_coef", "max_grad_norm",
               "learning_rate", "n_envs", "num_timesteps", "batch_size", "n_epochs", "normalize_advantage", "seed")


def read_sb3_zip(path: str) -> Tuple[Dict[str, torch.Tensor], Dict]:
    """(policy state dict, plain entries of `data`) of an SB3 zip.  Tensors are loaded with weights_only."""
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        if "policy.pth" not in names:
            raise ValueError(f"{path}: not a Stable-Baselines3 checkpoint (no policy.pth)")
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        data = json.loads(z.read("data").decode()) if "data" in names else {}
    plain = {k: data[k] for k in _PLAIN_KEYS if k in data and not isinstance(data[k], dict)}
    for k in ("observation_space", "action_space"):
        if isinstance(data.get(k), dict) and "_shape" in data[k]:
            plain[k + "_shape"] = tuple(data[k]["_shape"])
    return sd, plain


def load_sb3_policy(path: str, device="cpu", policy: Optional[ActorCritic] = None) -> Tuple[ActorCritic, Dict]:
    """Builds (or fills) an ActorCritic from an SB3 zip.  gSDE checkpoints (PPO_MPC default) are recognised by
    the [64, action_dim] log_std."""
    sd, plain = read_sb3_zip(path)
    obs_dim = sd["mlp_extractor.policy_net.0.weight"].shape[1]
    action_dim = sd["action_net.weight"].shape[0]
    use_sde = sd["log_std"].dim() == 2
    if sd["mlp_extractor.policy_net.0.weight"].shape[0] != 64 or "mlp_extractor.policy_net.4.weight" in sd:
        raise ValueError("only the reference's default net_arch (two 64-unit layers per tower) is supported")
    if policy is None:
        policy = ActorCritic(obs_dim, action_dim, use_sde=use_sde)
    policy.load_state_dict(sd, strict=True)
    return policy.to(device), plain


def save_sb3_policy(path: str, policy: ActorCritic, data: Optional[Dict] = None, optimizer: Optional[torch.optim.Optimizer] = None) -> None:
    def blob(obj) -> bytes:
        b = io.BytesIO()
        torch.save(obj, b)
        return b.getvalue()

    meta = {"use_sde": policy.use_sde, "policy_class": {":type:": "<class 'abc.ABCMeta'>", "__module__": "stable_baselines3.common.policies",
                                                         "repr": "ActorCriticPolicy (written by mpc_rl_for_avs_b200.checkpoint)"},
            "observation_space": {"_shape": [policy.mlp_extractor.policy_net[0].in_features]},
            "action_space": {"_shape": [policy.action_dim]}}
    meta.update({k: v for k, v in (data or {}).items() if isinstance(v, (int, float, bool, str, type(None)))})
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("data", json.dumps(meta, indent=1))
        z.writestr("policy.pth", blob({k: v.detach().cpu() for k, v in policy.state_dict().items()}))
        z.writestr("policy.optimizer.pth", blob(optimizer.state_dict() if optimizer is not None else {}))
        z.writestr("pytorch_variables.pth", blob(None))
        z.writestr("_stable_baselines3_version", "2.4.0")
        z.writestr("system_info.txt", f"- PyTorch: {torch.__version__}\n- writer: mpc_rl_for_avs_b200.checkpoint\n")
