import sys, os, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import helpers
import mpc_rl_for_avs_b200 as pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
obs, rs, has = pkg.make_scenarios(B, 8, seed=1234)
probs, _ = helpers.problems_from_obs(obs.numpy(), rs.numpy(), has.numpy(), w_distance=10.0, collision_check=True)
d = helpers.batch_from_problems(probs, 8)
lib = helpers.load_hostsim()
for mi in (60, 40, 30):
    cfg = helpers.hs_config(N=20, M=8, w_distance=10.0, max_iter=mi)
    r = helpers.hostsim_solve(lib, d, cfg)
    it, st = r["iters"], r["status"]
    print("max_iter", mi, "mean", it.mean(), "conv", (st == 0).mean(), "cap", ((st & 1) != 0).mean(), "stall", ((st & 16) != 0).mean(), "ls", ((st & 2) != 0).mean())
    print(" hist", np.histogram(it, bins=[0, 5, 10, 15, 20, 25, 30, 40, 50, 59, 61])[0] / B)
    np.save("/tmp/cost_%d.npy" % mi, r["cost"]); np.save("/tmp/st_%d.npy" % mi, st); np.save("/tmp/it_%d.npy" % mi, it)
c60, c30, c40 = np.load("/tmp/cost_60.npy"), np.load("/tmp/cost_30.npy"), np.load("/tmp/cost_40.npy")
for c, n in ((c40, 40), (c30, 30)):
    rel = (c - c60) / (1 + np.abs(c60))
    print("cap", n, "rel cost worse >1e-4:", (rel > 1e-4).mean(), ">1e-2:", (rel > 1e-2).mean())
