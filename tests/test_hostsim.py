"""CPU checks of the SOLVER LOGIC: the kernel's per-problem device functions (mpc_core.cuh) are
compiled for the host by the test-only harness (tests/hostsim) and compared with the oracle, in
float (the device arithmetic) and double.  These mirror the -m gpu parity tests so that the
algorithm is covered where no GPU exists; they do not exercise the CUDA library."""
import numpy as np
import pytest

import helpers
from helpers import orc


def _golden_batch(g):
    return {k[len("batch_"):]: v for k, v in g.items() if k.startswith("batch_")}


def _problems(g):
    probs, _ = helpers.problems_from_obs(g["obs"], g["ref_speed"], g["has_ref_speed"], w_distance=float(g["w_distance"]),
                                         collision_check=bool(g["collision_check"]))
    return probs


@pytest.mark.parametrize("name,M,wd", [("golden_track", 0, 0.0), ("golden_coll", 8, 10.0)])
@pytest.mark.parametrize("use_double", [False, True])
def test_rollout_cost_matches_oracle(hostsim, name, M, wd, use_double):
    g = helpers.load_golden(name)
    probs = _problems(g)[:64]
    d = {k: (v[..., :64] if v.ndim > 1 else v[:64]) for k, v in _golden_batch(g).items()}
    d = {k: np.ascontiguousarray(v) for k, v in d.items()}
    rng = np.random.default_rng(3)
    U = np.stack([rng.uniform(-5, 5, (64, 20)), rng.uniform(-1.0, 1.0, (64, 20))], axis=-1).astype(np.float32)
    U[:16] = g["oracle_U"][:16].astype(np.float32)
    X, c6, tot = helpers.hostsim_rollout_cost(hostsim, d, helpers.hs_config(M=M, w_distance=wd), U, use_double)
    for i in range(64):
        p = helpers.problem_f32(probs[i])
        Xo = orc.rollout(p.s0, U[i].astype(np.float64), p.dt)
        co = orc.cost_components(Xo, U[i].astype(np.float64), p)
        assert np.max(np.abs(X[i] - Xo) / np.maximum(np.abs(Xo), 1.0)) <= 1e-5, i
        near_disc, dist_tol = helpers.distance_conditioning(p, U[i])
        if near_disc:
            continue
        tol = 1e-5 * np.maximum(np.abs(co), 1.0)
        tol[4] += dist_tol
        assert np.all(np.abs(c6[i] - co) <= tol), (i, c6[i], co)
        to = orc.total_cost_from_components(co, p)
        assert abs(tot[i] - to) <= 1e-5 * max(abs(to), 1.0) + p.w_distance * dist_tol


@pytest.mark.parametrize("name,M,wd", [("golden_track", 0, 0.0), ("golden_coll", 8, 10.0), ("golden_holdout", 8, 10.0),
                                       ("golden_holdout_1k", 8, 10.0)])
@pytest.mark.parametrize("n_starts", [1, 4])
def test_solver_logic_against_golden(hostsim, name, M, wd, n_starts):
    """Host build of the device code against the best known optimum of the CPU portfolio (IPOPT-like interior point on
    the literal multiple-shooting NLP + SLSQP, oracle/ipm_oracle.py): the bars of helpers.PARITY_BARS."""
    g = helpers.load_golden(name)
    probs = _problems(g)
    B = len(probs)
    r = helpers.hostsim_solve_init(hostsim, _golden_batch(g), helpers.hs_config(M=M, w_distance=wd), n_starts=n_starts)
    st = helpers.solve_parity_stats(r, g, probs)
    helpers.assert_parity_bars(st, name, n_starts)
    # same first control almost always means the same optimum (a shared pinned first control with a
    # different tail is the exception): then the costs agree
    rel = np.abs(st["cost64"] - g["oracle_cost"]) / np.maximum(np.abs(g["oracle_cost"]), 1.0)
    assert np.mean(rel[st["same_mask"] & st["conv_mask"]] <= 1e-4) >= 0.9
    # every converged (status 0) solution is confirmed by the oracle started at it (sample of 32)
    for i in np.nonzero(st["conv_mask"])[0][:32]:
        ok, du0, gain = helpers.oracle_warm_confirms(probs[i], r["U"][i])
        assert ok, (i, du0, gain)
    # every iterate respects the reference's bounds, converged or not
    for i in range(B):
        X = orc.rollout(probs[i].s0, r["U"][i].astype(np.float64))
        assert X[1:, 3].min() >= -1e-4 and X[1:, 3].max() <= 30 + 1e-4 and np.abs(X[1:, 2]).max() <= np.pi + 1e-4


def test_float_and_double_agree_on_converged_problems(hostsim):
    g = helpers.load_golden("golden_track")
    d = _golden_batch(g)
    cfg = helpers.hs_config(M=0)
    rf = helpers.hostsim_solve(hostsim, d, cfg, use_double=False)
    rd = helpers.hostsim_solve(hostsim, d, cfg, use_double=True)
    both = (rf["status"] == 0) & (rd["status"] == 0)
    du = np.max(np.abs(rf["actions"] - rd["actions"]), axis=1)
    assert both.mean() > 0.85
    assert np.mean(du[both] <= 1e-3) >= 0.95      # the rest are different basins reached through rounding


def test_survey_known_answers(hostsim):
    ref = helpers.REF
    cases = [((2, 45, -np.pi / 2, 8), 125.78764721, (5.0, 0.0)),
             ((3, 30, -np.pi / 2 + 0.1, 5), 3143.38671483, (5.0, -0.62521401)),
             ((ref[48, 0] + 0.3, ref[48, 1] - 0.2, ref[48, 3] + 0.05, 9), 31.61840094, (5.0, -0.68614018))]
    probs = []
    for s0, _, _ in cases:
        s0 = np.array(s0, float)
        idx = orc.nearest_index(s0[:2], ref[:, :2])
        probs.append(orc.Problem(s0=s0, ego_index=idx, ref_v=ref[np.minimum(idx + np.arange(20), 84), 2].copy()))
    r = helpers.hostsim_solve(hostsim, helpers.batch_from_problems(probs, 0), helpers.hs_config(M=0))
    for i, (_, f, u0) in enumerate(cases):
        assert np.max(np.abs(r["actions"][i] - np.array(u0))) <= 1e-3
        assert abs(r["cost"][i] - f) <= 1e-4 * f


def test_horizon16_solver_logic(hostsim):
    """Shipped horizon (config/cfg.yaml:90) through the same device functions."""
    import mpc_rl_for_avs_b200 as pkg
    B, M, N = 32, 9, 16
    obs, rs, has = pkg.make_scenarios(B, M, seed=31)
    probs, _ = helpers.problems_from_obs(obs.numpy(), rs.numpy(), has.numpy(), w_distance=10.0, collision_check=True, N=N)
    r = helpers.hostsim_solve(hostsim, helpers.batch_from_problems(probs, M), helpers.hs_config(N=N, M=M, w_distance=10.0))
    conv = r["status"] == 0
    assert conv.mean() >= 0.8
    for i in np.nonzero(conv)[0][:12]:
        ok, du0, gain = helpers.oracle_warm_confirms(probs[i], r["U"][i])
        assert ok, (i, du0, gain)
