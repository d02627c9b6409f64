"""Small end-to-end case for compute-sanitizer (memcheck / racecheck), one tool per gpurun call:
    python tools/sanitize_case.py && compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpc_rl_for_avs_b200 as pkg

for M, N, B in ((8, 20, 700), (9, 16, 333), (0, 20, 257)):
    obs, rs, has = pkg.make_scenarios(B, M, seed=5)
    agent = pkg.BatchedPureMPC({"horizon": N}, vehicles_count=M + 1, max_batch=B, collision_check=M > 0, weight_distance=10.0 if M else 0.0)
    rsn = torch.where(has.reshape(-1, 1), rs, torch.full_like(rs, float("nan"))).cuda()
    for _ in range(2):
        a, U = agent.predict_batch(obs.cuda(), ref_speed=rsn, return_controls=True)
    ws = agent.workspace(B)
    X, c6, tot = agent.rollout_cost(ws, U)
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and torch.isfinite(tot).all()
    agent.close()
print("sanitize case ok")
