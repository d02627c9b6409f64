"""Generates the committed golden fixtures from the FP64 oracle (oracle/mpc_oracle.py).

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

The reference itself cannot be imported here (casadi / shapely / highway-env are not installable,
SURVEY 8-c), so these vectors hold the ORACLE's solutions, not CasADi/IPOPT's; the oracle itself is pinned to
the reference's own code by make_reference_golden.py / golden_reference.npz.  Files:
  golden_track.npz   256 scenarios, M=0, tracking objective (BASELINE config 2 type)
  golden_coll.npz    256 scenarios, M=8, collision check + regeneration + distance cost 10 (config 3 type)
  golden_holdout.npz 256 more of the config-3 type from another seed: the device solver's start portfolio was selected on
                     the first two sets, this one only measures
  golden_holdout_1k.npz  1024 more (seed 777), generated last: the out-of-sample rates quoted in DESIGN.md come from here
Each holds the observations (float32), the parsed problem descriptors, collision outputs and the
oracle's NLP solution from the reference's cold start.
"""
import os
import sys
import multiprocessing as mp

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import helpers  # noqa: E402
from helpers import orc  # noqa: E402
import mpc_rl_for_avs_b200 as pkg  # noqa: E402


def _solve(p):
    import ipm_oracle as ipm
    b = ipm.best_known_optimum(p)
    X = orc.rollout(p.s0, b["U"], p.dt)
    b["components"] = orc.cost_components(X, b["U"], p)
    b["kkt"] = orc.kkt_residual(b["U"], p)[0]
    return b


def make(name, B, M, seed, w_distance, collision_check):
    obs, rs, has = pkg.make_scenarios(B, M, seed=seed)
    obs, rs, has = obs.numpy(), rs.numpy(), has.numpy()
    probs, agents = helpers.problems_from_obs(obs, rs, has, w_distance=w_distance, collision_check=collision_check)
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        sols = pool.map(_solve, probs, chunksize=4)
    d = helpers.batch_from_problems(probs, M)
    out = dict(obs=obs, ref_speed=rs, has_ref_speed=has, w_distance=np.float64(w_distance),
               collision_check=np.bool_(collision_check), n_obstacles=np.int64(M))
    out.update({"batch_" + k: v for k, v in d.items()})
    # oracle_* = best confirmed optimum of the CPU portfolio (ipm_oracle.best_known_optimum: interior point on the literal
    # NLP + SLSQP from the reference's cold start, from three steering-pulse starts and from two path-following starts:
    # seven runs per problem, every result pulled back into the feasible set); ipm_* = the
    # IPOPT-like interior point on the literal multiple-shooting NLP alone (basin predictor, polished);
    # in_path = the horizon stays on the 85-point path (ego_index + N <= 84): past it the reference point is
    # frozen at the path end while the reference speed is not, and the NLP is ill-posed
    out["oracle_U"] = np.array([s["U"] for s in sols])
    out["oracle_cost"] = np.array([s["cost"] for s in sols])
    out["oracle_components"] = np.array([s["components"] for s in sols])
    out["oracle_success"] = np.array([s["success"] for s in sols])
    out["oracle_source"] = np.array([s["source"] for s in sols])
    out["oracle_kkt"] = np.array([s["kkt"] for s in sols])
    out["ipm_U"] = np.array([s["ipm_U"] for s in sols])
    out["ipm_cost"] = np.array([s["ipm_cost"] for s in sols])
    out["ipm_confirmed"] = np.array([s["ipm_confirmed"] for s in sols])
    out["ipm_raw_U"] = np.array([s["ipm_raw_U"] for s in sols])
    out["ipm_status"] = np.array([s["ipm_status"] for s in sols])
    out["ipm_restoration"] = np.array([s["ipm_restoration"] for s in sols])
    out["ipm_iters"] = np.array([s["ipm_iters"] for s in sols])
    out["slsqp_U"] = np.array([s["slsqp_U"] for s in sols])
    out["slsqp_cost"] = np.array([s["slsqp_cost"] for s in sols])
    out["slsqp_confirmed"] = np.array([s["slsqp_confirmed"] for s in sols])
    out["in_path"] = np.array([p.ego_index + p.N <= 84 for p in probs])
    Mx = max(M, 1)
    flags = np.zeros((B, Mx), np.uint8); cidx = -np.ones((B, Mx), np.int32); deg = np.zeros(B, np.uint8)
    if collision_check:
        for i in range(B):
            parsed = orc.parse_obs(obs[i], obs.shape[1])
            res = orc.detect_collisions(parsed.ego, parsed.others)
            deg[i] = res.degenerate
            for m, (f, c) in enumerate(zip(res.agent_collide, res.conflict_index)):
                flags[i, m] = f
                cidx[i, m] = -1 if c is None else c
    out["agent_collide"], out["conflict_index"], out["degenerate"] = flags, cidx, deg
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written:", B, "problems; oracle success", out["oracle_success"].mean(), "source ipm", float(np.mean(out["oracle_source"] == "ipm")),
          "ipm confirmed", out["ipm_confirmed"].mean(), "ipm restoration", out["ipm_restoration"].mean(), "in_path", out["in_path"].mean(),
          "collide frac", float(np.mean(d["is_collide"])), "degenerate", int(deg.sum()))


SETS = {"golden_track": (256, 0, 11, 0.0, False), "golden_coll": (256, 8, 12, 10.0, True),
        # hold-out set: never looked at while the device solver's start portfolio and its thresholds were chosen
        "golden_holdout": (256, 8, 2024, 10.0, True),
        # the same, four times larger (rates to +-1 %): generated last, never used for a decision
        "golden_holdout_1k": (1024, 8, 777, 10.0, True)}

if __name__ == "__main__":
    for name in (sys.argv[1:] or SETS):
        make(name, *SETS[name])
