"""Solve-parity report on the golden sets with the HOST BUILD of the device code (tests/hostsim; test harness, not the
product): converged rate, J_gpu <= J_oracle rate, first-control agreement with the best known optimum and with the
IPOPT-like oracle, overall and for the problems whose horizon stays on the reference path (`in_path`).

    python tools/solve_parity_report.py [--double] [--max-iter N]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import helpers  # noqa: E402
from helpers import orc  # noqa: E402


def report(name, r, g, probs, verbose=False):
    B = len(probs)
    gc = np.array([orc.objective(r["U"][i].astype(np.float64), probs[i]) for i in range(B)])
    st = r["status"]
    conv = st == 0
    oc, ic = g["oracle_cost"], g["ipm_cost"]
    below = gc <= oc * (1 + 1e-6) + 1e-6
    below_ipm = gc <= ic * (1 + 1e-6) + 1e-6
    same = np.max(np.abs(r["actions"] - g["oracle_U"][:, 0, :]), axis=1) <= 1e-3
    same_ipm = np.max(np.abs(r["actions"] - g["ipm_U"][:, 0, :]), axis=1) <= 1e-3
    clean = g["ipm_confirmed"] & ~g["ipm_restoration"]
    ip = g["in_path"]
    out = {}
    for tag, m in (("all", np.ones(B, bool)), ("in_path", ip), ("off_path", ~ip)):
        out[tag] = dict(n=int(m.sum()), conv=conv[m].mean(), below=below[m].mean(), same=same[m].mean(),
                        below_ipm=below_ipm[m].mean(), same_ipm=same_ipm[m].mean(),
                        same_ipm_clean=same_ipm[m & clean].mean() if (m & clean).any() else float("nan"))
        o = out[tag]
        print(f"{name:13s} {tag:8s} n {o['n']:4d} conv {o['conv']:.3f} | vs best: below {o['below']:.3f} same-u0 {o['same']:.3f} | "
              f"vs ipm: below {o['below_ipm']:.3f} same-u0 {o['same_ipm']:.3f} (clean ipm {o['same_ipm_clean']:.3f})")
    it = r["iters"]
    print(f"{'':13s} iters mean {it.mean():.1f} p50 {np.median(it):.0f} p90 {np.percentile(it, 90):.0f} p99 {np.percentile(it, 99):.0f} max {it.max()} | "
          f"status counts { {int(s): int((st == s).sum()) for s in np.unique(st)} }")
    rel = (gc - oc) / np.maximum(1.0, np.abs(oc))
    print(f"{'':13s} excess cost over best where above: n {(~below).sum()} median {np.median(rel[~below]) if (~below).any() else 0:.2e} "
          f"p90 {np.percentile(rel[~below], 90) if (~below).any() else 0:.2e}")
    if verbose:
        for i in np.nonzero(~below | ~conv)[0]:
            p = probs[i]
            print(f"   {i:3d} idx {p.ego_index:2d} col {int(p.is_collide)} st {st[i]:2d} it {it[i]:3d} J {gc[i]:.7g} best {oc[i]:.7g} ({g['oracle_source'][i]}) ipm {ic[i]:.7g} "
                  f"u0 {r['actions'][i].round(4)} best {g['oracle_U'][i, 0].round(4)}")
    return out, gc


def run(name, use_double=False, verbose=False, **cfgkw):
    g = helpers.load_golden(name)
    M = int(g["n_obstacles"]); wd = float(g["w_distance"])
    probs, _ = helpers.problems_from_obs(g["obs"], g["ref_speed"], g["has_ref_speed"], w_distance=wd,
                                         collision_check=bool(g["collision_check"]))
    d = helpers.batch_from_problems(probs, M)
    lib = helpers.load_hostsim()
    r = helpers.hostsim_solve(lib, d, helpers.hs_config(N=20, M=max(M, 0), w_distance=wd, **cfgkw), use_double=use_double)
    return report(name, r, g, probs, verbose)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--double", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--max-iter", type=int, default=None)
    a = ap.parse_args()
    kw = {} if a.max_iter is None else {"max_iter": a.max_iter}
    for nm in ("golden_track", "golden_coll"):
        run(nm, a.double, a.verbose, **kw)
