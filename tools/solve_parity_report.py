"""Solve-parity report on the golden sets with the HOST BUILD of the device code (tests/hostsim; test harness, not the
product): converged / settled rates, J_gpu <= J_oracle rate, first-control agreement with the best known optimum and
with the IPOPT-like oracle, overall and for the problems whose horizon stays on the reference path (`in_path`).

    python tools/solve_parity_report.py [--double] [--n-starts 1 4] [--verbose]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import helpers  # noqa: E402


def report(name, r, g, probs, verbose=False):
    s = helpers.solve_parity_stats(r, g, probs)
    for tag in ("all", "in_path"):
        o = s[tag]
        print(f"{name:13s} {tag:8s} n {o['n']:4d} conv {o['conv']:.3f} settled {o['settled']:.3f} | vs best: below {o['below']:.3f} within 1% {o['near']:.3f} "
              f"same-u0 {o['same']:.3f} | vs ipm: same-u0 {o['same_ipm']:.3f}")
    it, st = r["iters"], r["status"]
    print(f"{'':13s} iters mean {it.mean():.1f} p50 {np.median(it):.0f} p90 {np.percentile(it, 90):.0f} p99 {np.percentile(it, 99):.0f} max {it.max()} | "
          f"status counts { {int(x): int((st == x).sum()) for x in np.unique(st)} }")
    gc, oc, below = s["cost64"], g["oracle_cost"], s["below_mask"]
    rel = (gc - oc) / np.maximum(1.0, np.abs(oc))
    if (~below).any():
        print(f"{'':13s} excess cost over best where above: n {(~below).sum()} median {np.median(rel[~below]):.2e} p90 {np.percentile(rel[~below], 90):.2e}")
    if verbose:
        for i in np.nonzero(~below | (st != 0))[0]:
            p = probs[i]
            print(f"   {i:3d} idx {p.ego_index:2d} col {int(p.is_collide)} st {st[i]:2d} it {it[i]:3d} J {gc[i]:.7g} best {oc[i]:.7g} ({g['oracle_source'][i]}) "
                  f"ipm {g['ipm_cost'][i]:.7g} u0 {r['actions'][i].round(4)} best {g['oracle_U'][i, 0].round(4)}")
    return s


def run(name, use_double=False, verbose=False, n_starts=1, **cfgkw):
    g = helpers.load_golden(name)
    M = int(g["n_obstacles"]); wd = float(g["w_distance"])
    probs, _ = helpers.problems_from_obs(g["obs"], g["ref_speed"], g["has_ref_speed"], w_distance=wd,
                                         collision_check=bool(g["collision_check"]))
    d = helpers.batch_from_problems(probs, M)
    lib = helpers.load_hostsim()
    r = helpers.hostsim_solve_init(lib, d, helpers.hs_config(N=20, M=max(M, 0), w_distance=wd, **cfgkw), use_double=use_double,
                                   n_starts=n_starts)
    return report(name, r, g, probs, verbose)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--double", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--n-starts", type=int, nargs="+", default=[1, 4])
    a = ap.parse_args()
    for S in a.n_starts:
        print(f"==== n_starts {S}")
        for nm in ("golden_track", "golden_coll", "golden_holdout", "golden_holdout_1k"):
            run(nm, a.double, a.verbose, n_starts=S)
