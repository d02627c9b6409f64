import os, sys, subprocess, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
if len(sys.argv) > 1:
    import mpc_rl_for_avs_b200 as pkg
    B, M = 65536, 8
    obs, rs, has = pkg.make_scenarios(B, M, seed=1234)
    rs_dev = torch.where(has.reshape(-1, 1), rs, torch.full_like(rs, float("nan"))).cuda()
    agent = pkg.BatchedPureMPC({"horizon": 20, "weight_speed": 1.0, "weight_control": 1.0, "weight_input_diff": 1.0},
                               vehicles_count=M + 1, max_batch=B, collision_check=True, weight_distance=10.0)
    a = agent.predict_batch(obs.cuda(), ref_speed=rs_dev)
    torch.cuda.synchronize()
    np.savez(sys.argv[1], a=a.cpu().numpy(), it=agent.iters[:B].cpu().numpy(), st=agent.status[:B].cpu().numpy(), c=agent.cost[:B].cpu().numpy())
else:
    for name, lib in (("/tmp/c1.npz", "libmpcb200.so"), ("/tmp/c0.npz", "libmpc_nocompact.so")):
        subprocess.check_call([sys.executable, __file__, name], env=dict(os.environ, MPC_LIB_NAME=lib))
    x, y = np.load("/tmp/c1.npz"), np.load("/tmp/c0.npz")
    d = np.where((x["a"] != y["a"]).any(1) | (x["it"] != y["it"]) | (x["st"] != y["st"]))[0]
    print("differing problems", len(d), "of", len(x["it"]))
    print("iters of differing (compact / not):", x["it"][d][:20], y["it"][d][:20])
    print("status:", x["st"][d][:20], y["st"][d][:20])
    print("cost:", x["c"][d][:10], y["c"][d][:10])
    print("min iters among differing (no-compact run):", y["it"][d].min() if len(d) else None, "hist of iters (no-compact) for differing", np.bincount(y["it"][d])[:70] if len(d) else None)
    print("idx", d[:30])
