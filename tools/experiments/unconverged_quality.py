"""How good are the iterates the solver returns WITHOUT the convergence flag (stalled / iteration cap)?
Host build of the device code on the golden collision set; each returned U is handed to the oracle's SLSQP
as a warm start: du0 = how far the first control still moves, gain = relative cost still to be gained."""
import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import helpers
from helpers import orc
g = helpers.load_golden(sys.argv[1] if len(sys.argv) > 1 else "golden_coll")
M = int(g["n_obstacles"]); wd = float(g["w_distance"])
probs, _ = helpers.problems_from_obs(g["obs"], g["ref_speed"], g["has_ref_speed"], w_distance=wd, collision_check=bool(g["collision_check"]))
d = helpers.batch_from_problems(probs, M)
lib = helpers.load_hostsim()
r = helpers.hostsim_solve(lib, d, helpers.hs_config(N=20, M=max(M, 0), w_distance=wd))
st = r["status"]
print("n", len(st), "converged", (st == 0).mean(), "stalled", ((st & 16) != 0).mean(), "cap", ((st & 1) != 0).mean(), "ls-fail", ((st & 2) != 0).mean())
rows = []
for i in np.where(st != 0)[0]:
    ok, du0, gain = helpers.oracle_warm_confirms(probs[i], r["U"][i].astype(np.float64))
    rows.append((int(st[i]), int(r["iters"][i]), du0, gain))
rows = np.array(rows)
for code, name in ((16, "stalled"), (1, "cap"), (2, "ls-fail")):
    m = (rows[:, 0].astype(int) & code) != 0
    if m.any():
        print(name, m.sum(), "du0 median %.2e p90 %.2e max %.2e | rel gain median %.2e p90 %.2e max %.2e" % (
            np.median(rows[m, 2]), np.quantile(rows[m, 2], 0.9), rows[m, 2].max(), np.median(rows[m, 3]), np.quantile(rows[m, 3], 0.9), rows[m, 3].max()))
print("all unconverged: du0<1e-3: %.2f  du0<1e-2: %.2f  gain<1e-4: %.2f  gain<1e-3: %.2f" % (
    (rows[:, 2] < 1e-3).mean(), (rows[:, 2] < 1e-2).mean(), (rows[:, 3] < 1e-4).mean(), (rows[:, 3] < 1e-3).mean()))
