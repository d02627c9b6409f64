"""Diagnostic: CUDA solve of the hold-out golden set, results dumped for offline analysis with the host harness."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
import mpc_rl_for_avs_b200 as pkg
name = sys.argv[1] if len(sys.argv) > 1 else "golden_holdout"
g = helpers.load_golden(name)
M = int(g["n_obstacles"]); wd = float(g["w_distance"])
batch = {k[6:]: g[k] for k in g if k.startswith("batch_")}
B = batch["ego_index"].shape[0]
out = {}
for S in (1, 4, 8):
    agent = pkg.BatchedPureMPC({"horizon": 20, "weight_speed": 1, "weight_control": 1, "weight_input_diff": 1}, vehicles_count=M + 1, max_batch=B,
                               collision_check=False, weight_distance=wd, n_starts=S)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in batch.items()}
    actions, U = agent.solve_batch(dev, return_controls=True)
    torch.cuda.synchronize()
    out[f"U{S}"] = U.cpu().numpy(); out[f"st{S}"] = agent.status[:B].cpu().numpy(); out[f"it{S}"] = agent.iters[:B].cpu().numpy()
    out[f"cost{S}"] = agent.cost[:B].cpu().numpy()
np.savez(os.path.join(ROOT, "gpurun_out", f"dump_{name}.npz"), **out)
print("dumped", {k: v.shape for k, v in out.items()})
