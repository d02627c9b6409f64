"""Large-sample check of the status-0 certificate on the bench workload: every solve returned with status 0 is handed to
the oracle's NLP solver (tests/helpers.py: oracle_warm_confirms); the unconfirmed ones are printed and dumped.

    python tools/experiments/confirm_sample.py [problems] [n_starts]
"""
import multiprocessing as mp
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

H, M, W_DIST = 20, 8, 10.0


def _check(item):
    import mpc_oracle as orc
    import helpers
    obs, r, U = item
    ag = orc.OraclePureMPCAgent(horizon=H, vehicles_count=M + 1, weight_distance=W_DIST, collision_check=True)
    parsed = orc.parse_obs(obs, M + 1)
    ag.check_collision(parsed)
    prob = ag.build_problem(parsed, None, None if r is None else np.asarray(r).reshape(1, 1))
    ok, du0, gain = helpers.oracle_warm_confirms(prob, U)
    return ok, du0, gain, orc.objective(np.asarray(U, np.float64), prob), prob.ego_index, int(prob.is_collide)


if __name__ == "__main__":
    import torch
    import mpc_rl_for_avs_b200 as pkg
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    obs, rs, has = pkg.make_scenarios(n, M, seed=1234)
    agent = pkg.BatchedPureMPC({"horizon": H, "weight_speed": 1, "weight_control": 1, "weight_input_diff": 1}, vehicles_count=M + 1,
                               max_batch=n, collision_check=True, weight_distance=W_DIST, n_starts=S)
    rsn = torch.where(has.reshape(-1), rs.reshape(-1), torch.full((n,), float("nan"))).float()
    act, U = agent.predict_batch(obs.cuda(), ref_speed=rsn.cuda(), return_controls=True)
    torch.cuda.synchronize()
    U, st = U.cpu().numpy(), agent.status[:n].cpu().numpy()
    obs, rs, has = obs.numpy(), rs.numpy(), has.numpy()
    idx = np.nonzero(st == 0)[0]
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(_check, [(obs[i], (rs[i] if has[i] else None), U[i]) for i in idx], chunksize=4)
    bad = [(int(i),) + r[1:] for i, r in zip(idx, res) if not r[0]]
    print(f"status 0: {len(idx)} of {n}; confirmed {len(idx) - len(bad)}; unconfirmed {len(bad)}")
    for b in bad:
        print("  problem %d du0 %.3g rel gain %.3g J %.8g ego_index %d collide %d" % b)
    np.savez(os.path.join(ROOT, "gpurun_out", "confirm_sample.npz"), bad=np.array([b[0] for b in bad], int), U=U, status=st, obs=obs, rs=rs, has=has)
