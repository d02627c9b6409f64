"""Greedy choice of the start portfolio on a tuning set, checked on the hold-out set (host build of the device code).

    python tools/experiments/start_selection.py <tuning.npz> <holdout.npz>
Every candidate start is solved alone (U0 = the start's controls); a portfolio's result for a problem is the candidate with the
lowest FP32 objective (what k_select does); score = share of problems with J <= J_oracle (1 + 1e-5) + 1e-4.
Candidates: 25 constant-acceleration / steering-pulse starts and 8 path-following starts (numpy prototype of
mpc_core.cuh: apply_start, kind kStartPath).  The tuning set of round 2 was `make_golden.make("golden_dev2", 1024, 8, 31337,
10.0, True)` (not committed); the hold-out set is tests/golden/golden_holdout_1k.npz.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import helpers

CANDS = [(0, 0, 0), (-5, 0, 0), (0, -0.9, 3), (0, 0.4, 3), (5, 0.9, 3), (0, -0.4, 3), (0, -0.4, 99), (5, -0.4, 3),      # 0-7: the shipped table
         (5, 0, 0), (0, 0.9, 3), (-5, -0.9, 3), (-5, 0.9, 3), (0, 0.4, 99), (0, 0.9, 99), (0, -0.9, 99), (2.5, 0, 0), (-2.5, 0, 0),
         (0, 0.2, 3), (0, -0.2, 3), (5, 0.4, 3), (-5, 0.4, 3), (-5, -0.4, 3), (0, 0.9, 6), (0, -0.9, 6), (5, -0.9, 3)]


PATH_CANDS = [("ref", 1), ("hold", 1), ("brake", 1), ("accel", 1), ("ref", 2), ("ref", 3), ("hold", 2), ("accel", 2)]
SB_MAX = float(np.sin(np.arctan(0.5 * np.tan(np.pi / 3))))


def path_following_controls(p, amode, look):
    """Roll the model forward steering at the path point `look` rows ahead, accelerating as `amode` says."""
    import mpc_oracle as orc
    REF = helpers.REF
    U = np.zeros((p.N, 2)); s = p.s0.copy()
    for k in range(p.N):
        j1 = min(p.ego_index + k + look, 84)
        vref = p.ref_v[min(k, len(p.ref_v) - 1)]
        a = {"ref": (vref - s[3]) / p.dt, "hold": 0.0, "brake": -5.0, "accel": 5.0}[amode]
        a = np.clip(np.clip(a, -5, 5), (0 - s[3]) / p.dt, (30 - s[3]) / p.dt)
        tx, ty = REF[j1, 0], REF[j1, 1]
        des = np.arctan2(ty - s[1], tx - s[0]) if np.hypot(tx - s[0], ty - s[1]) > 1e-3 else REF[j1, 3]
        dth = np.arctan2(np.sin(des - s[2]), np.cos(des - s[2]))
        sb = np.clip(dth / (p.dt * max(s[3], 1e-3) / 2.5), -SB_MAX, SB_MAX)
        U[k] = (a, np.arcsin(np.clip(sb / np.sqrt(0.25 + 0.75 * sb * sb), -1, 1)))
        s = orc.step(s, U[k], p.dt)
    return orc.repair_feasible(U, p)


def solve_all(path):
    g = dict(np.load(path))
    M, wd = int(g["n_obstacles"]), float(g["w_distance"])
    probs, _ = helpers.problems_from_obs(g["obs"], g["ref_speed"], g["has_ref_speed"], w_distance=wd, collision_check=True)
    d = helpers.batch_from_problems(probs, M)
    lib = helpers.load_hostsim()
    B = len(probs)
    J = np.zeros((len(CANDS) + len(PATH_CANDS), B))
    for c, (a, dl, nk) in enumerate(CANDS):
        U0 = np.zeros((B, 20, 2), np.float32); U0[:, :, 0] = a; U0[:, :min(nk, 20), 1] = dl
        J[c] = helpers.hostsim_solve_init(lib, d, helpers.hs_config(N=20, M=M, w_distance=wd), U0=U0, n_starts=1)["cost"]
    for c, (am, lk) in enumerate(PATH_CANDS):
        U0 = np.stack([path_following_controls(p, am, lk) for p in probs]).astype(np.float32)
        J[len(CANDS) + c] = helpers.hostsim_solve_init(lib, d, helpers.hs_config(N=20, M=M, w_distance=wd), U0=U0, n_starts=1)["cost"]
    return J, g["oracle_cost"], g["in_path"].astype(bool)


def score(J, oc, sel):
    best = J[sel].min(axis=0)
    return best <= oc * (1 + 1e-5) + 1e-4


if __name__ == "__main__":
    Jt, oct_, ipt = solve_all(sys.argv[1])
    Jh, och, iph = solve_all(sys.argv[2])
    NAMES = [str(c) for c in CANDS] + ["path:%s:%d" % v for v in PATH_CANDS]
    shipped = [0, 1, 2, 3]
    print("first pulse table, 4 starts: tuning %.3f (in-path %.3f)  hold-out %.3f (in-path %.3f)" % (score(Jt, oct_, shipped).mean(), score(Jt, oct_, shipped)[ipt].mean(),
                                                                                         score(Jh, och, shipped).mean(), score(Jh, och, shipped)[iph].mean()))
    print("first pulse table, 8 starts: tuning %.3f  hold-out %.3f" % (score(Jt, oct_, list(range(8))).mean(), score(Jh, och, list(range(8))).mean()))
    sel = [0]
    for _ in range(7):
        gains = [(score(Jt, oct_, sel + [c]).mean(), c) for c in range(len(NAMES)) if c not in sel]
        sc, c = max(gains)
        sel.append(c)
        print("greedy +%-16s tuning %.3f (in-path %.3f)  hold-out %.3f (in-path %.3f)" % (NAMES[c], sc, score(Jt, oct_, sel)[ipt].mean(), score(Jh, och, sel).mean(),
                                                                                          score(Jh, och, sel)[iph].mean()))
    single = [(round(float(score(Jt, oct_, [c]).mean()), 3), NAMES[c]) for c in range(len(NAMES))]
    print("single-start scores on the tuning set:", sorted(single, reverse=True)[:8])
