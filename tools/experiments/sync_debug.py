import sys, traceback
sys.path.insert(0, "/root/repo")
import torch
import mpc_rl_for_avs_b200 as pkg
from mpc_rl_for_avs_b200.rl import A2CMPC, PPOMPC, BatchedIntersectionEnv
B = 512
cfg = {"horizon": 16, "weight_speed": 1, "weight_control": 1, "weight_input_diff": 1}
for Algo in (A2CMPC, PPOMPC):
    mpc = pkg.BatchedPureMPC(cfg, vehicles_count=10, max_batch=B, collision_check=True)
    algo = Algo(BatchedIntersectionEnv(B, 9, device="cuda", seed=9, duration_steps=20), mpc, n_steps=4, graph=False)
    algo.policy.reset_noise(B)
    algo._transition()
    algo._row.zero_()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        algo._transition()
        algo._row.zero_()
        print(Algo.__name__, "transition is sync-free")
    except Exception:
        traceback.print_exc()
    torch.cuda.set_sync_debug_mode("default")
    algo.graph = True
    try:
        algo.train_step(); algo.train_step()
        print(Algo.__name__, "graph ok", int(algo._row))
    except Exception:
        traceback.print_exc()
