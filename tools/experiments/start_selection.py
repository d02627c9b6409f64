"""Greedy choice of the start portfolio on a tuning set, checked on the hold-out set (host build of the device code).

    python tools/experiments/start_selection.py <tuning.npz> <holdout.npz>
Every candidate start is solved alone (U0 = the start's controls); a portfolio's result for a problem is the candidate with the
lowest FP32 objective (what k_select does); score = share of problems with J <= J_oracle (1 + 1e-5) + 1e-4.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import helpers

CANDS = [(0, 0, 0), (-5, 0, 0), (0, -0.9, 3), (0, 0.4, 3), (5, 0.9, 3), (0, -0.4, 3), (0, -0.4, 99), (5, -0.4, 3),      # 0-7: the shipped table
         (5, 0, 0), (0, 0.9, 3), (-5, -0.9, 3), (-5, 0.9, 3), (0, 0.4, 99), (0, 0.9, 99), (0, -0.9, 99), (2.5, 0, 0), (-2.5, 0, 0),
         (0, 0.2, 3), (0, -0.2, 3), (5, 0.4, 3), (-5, 0.4, 3), (-5, -0.4, 3), (0, 0.9, 6), (0, -0.9, 6), (5, -0.9, 3)]


def solve_all(path):
    g = dict(np.load(path))
    M, wd = int(g["n_obstacles"]), float(g["w_distance"])
    probs, _ = helpers.problems_from_obs(g["obs"], g["ref_speed"], g["has_ref_speed"], w_distance=wd, collision_check=True)
    d = helpers.batch_from_problems(probs, M)
    lib = helpers.load_hostsim()
    B = len(probs)
    J = np.zeros((len(CANDS), B))
    for c, (a, dl, nk) in enumerate(CANDS):
        U0 = np.zeros((B, 20, 2), np.float32); U0[:, :, 0] = a; U0[:, :min(nk, 20), 1] = dl
        J[c] = helpers.hostsim_solve_init(lib, d, helpers.hs_config(N=20, M=M, w_distance=wd), U0=U0, n_starts=1)["cost"]
    return J, g["oracle_cost"], g["in_path"].astype(bool)


def score(J, oc, sel):
    best = J[sel].min(axis=0)
    return best <= oc * (1 + 1e-5) + 1e-4


if __name__ == "__main__":
    Jt, oct_, ipt = solve_all(sys.argv[1])
    Jh, och, iph = solve_all(sys.argv[2])
    shipped = [0, 1, 2, 3]
    print("shipped 4 starts: tuning %.3f (in-path %.3f)  hold-out %.3f (in-path %.3f)" % (score(Jt, oct_, shipped).mean(), score(Jt, oct_, shipped)[ipt].mean(),
                                                                                         score(Jh, och, shipped).mean(), score(Jh, och, shipped)[iph].mean()))
    print("shipped 8 starts: tuning %.3f  hold-out %.3f" % (score(Jt, oct_, list(range(8))).mean(), score(Jh, och, list(range(8))).mean()))
    sel = [0]
    for _ in range(7):
        gains = [(score(Jt, oct_, sel + [c]).mean(), c) for c in range(len(CANDS)) if c not in sel]
        sc, c = max(gains)
        sel.append(c)
        print("greedy +%-14s tuning %.3f (in-path %.3f)  hold-out %.3f (in-path %.3f)" % (CANDS[c], sc, score(Jt, oct_, sel)[ipt].mean(), score(Jh, och, sel).mean(),
                                                                                          score(Jh, och, sel)[iph].mean()))
    single = [(score(Jt, oct_, [c]).mean(), CANDS[c]) for c in range(len(CANDS))]
    print("single-start scores on the tuning set:", sorted(single, reverse=True)[:8])
