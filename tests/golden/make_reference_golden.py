"""Pins the oracle against the REFERENCE'S OWN PYTHON, executed here.

    OMP_NUM_THREADS=1 python tests/golden/make_reference_golden.py        (needs /root/reference)

`agents/pure_mpc.py` imports casadi, shapely, gymnasium and matplotlib, none of which is installable
in this image.  This script puts small numeric stand-ins for those four packages into `sys.modules`
and then imports and runs the UNMODIFIED reference modules from /root/reference:

  * casadi  -> numbers instead of symbols.  `SX.sym('x', 4, N+1)` / `SX.sym('u', 2, N)` return the
    numeric trajectory this script injects; every expression the reference then builds (tracking /
    control / input-difference / final-state costs, the dynamics defects g, the bounds, the initial
    guess) is evaluated by the reference's own lines with numpy float64.  `nlpsol` records
    f, g, lbx, ubx, x0 and "returns" the injected point.  So the fixture holds the reference's
    OBJECTIVE, CONSTRAINT RESIDUALS, BOUNDS and COLD START at known trajectories; what it cannot hold
    is IPOPT's answer.
  * shapely -> `LineString.intersection` by the textbook parametric segment-segment formula
    (proper crossings only; collinear overlaps are reported as degenerate and such scenes are dropped).
    MultiPoint members are returned in lexicographic (x, y) order: an ASSUMPTION about GEOS, not a fact.
    Everything AROUND that primitive is the reference's code: nearest path index, the ego / other
    polylines, closest-time test, conflict index, the 10-step latch, reference-speed regeneration.
  * gymnasium / matplotlib -> empty shells (type annotation / unused plotting).

Outputs tests/golden/golden_reference.npz:
  single-step scenes (cold latch)   obs, RL reference speed, injected U and X, and what the reference
                                    computed: parsed ego state, ego_index, flags, conflict indices, memory,
                                    regenerated speed column, f, six components, max |g|, bounds, x0, action;
                                    plus the obstacle-distance component of agents/archive/pure_mpc.py:189-196
                                    (the formula BASELINE config 3 adds) evaluated by that file at the same point,
                                    and the literal objective of agents/pure_mpc_no_collision.py (BASELINE config 2)
  latch sequences                   the same agent object stepped through moving scenes
The reference is never copied: it is imported from where it lies and only its outputs are stored.
"""
import os
import sys
import types

os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

INJECT = {}      # name -> ndarray handed out by SX.sym
CAPTURE = {}     # what the reference passed to Function / nlpsol / the solver call


# ------------------------------------------------------------------------------------------- casadi stand-in
class _Vec(np.ndarray):
    def size1(self):
        return int(self.shape[0])

    def full(self):
        return np.asarray(self, dtype=np.float64).reshape(-1, 1)


def _vec(a):
    return np.asarray(a, dtype=np.float64).view(_Vec)


def _install_casadi():
    ca = types.ModuleType("casadi")

    class SX:
        @staticmethod
        def sym(name, r, c=1):
            v = np.array(INJECT[name], dtype=np.float64)
            assert v.shape == (r, c), (name, v.shape, (r, c))
            return v

    def vertcat(*args):
        return _vec(np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1, order="F") for a in args]))

    def reshape(x, r, c):
        return np.asarray(x, dtype=np.float64).reshape((r, c), order="F")      # casadi is column-major

    class Function:
        def __init__(self, name, ins, outs):
            CAPTURE["components"] = [float(o) for o in outs]

    class _Solver:
        def __init__(self, nlp):
            CAPTURE["f"] = float(nlp["f"])
            CAPTURE["g"] = np.asarray(nlp["g"], dtype=np.float64).copy()
            CAPTURE["x"] = np.asarray(nlp["x"], dtype=np.float64).copy()

        def __call__(self, x0, lbx, ubx, lbg, ubg):
            CAPTURE.update(x0=np.asarray(x0, dtype=np.float64), lbx=np.asarray(lbx, dtype=np.float64),
                           ubx=np.asarray(ubx, dtype=np.float64), lbg=np.asarray(lbg, dtype=np.float64),
                           ubg=np.asarray(ubg, dtype=np.float64))
            return {"x": _vec(CAPTURE["x"])}

        def stats(self):
            return {"success": True}

    ca.SX, ca.vertcat, ca.reshape, ca.Function = SX, vertcat, reshape, Function
    ca.nlpsol = lambda name, kind, nlp, opts: _Solver(nlp)
    ca.sin, ca.cos, ca.tan, ca.atan = np.sin, np.cos, np.tan, np.arctan
    ca.norm_2 = lambda v: np.sqrt(np.sum(np.asarray(v, dtype=np.float64) ** 2))
    ca.if_else = lambda c, a, b: a if c else b
    ca.sumsqr = lambda v: np.sum(np.asarray(v, dtype=np.float64) ** 2)
    ca.pi, ca.inf = np.pi, np.inf
    sys.modules["casadi"] = ca


# ------------------------------------------------------------------------------------------- shapely stand-in
class GEOSException(Exception):
    pass


class _Point:
    geom_type = "Point"
    is_empty = False

    def __init__(self, x, y):
        self.x, self.y = x, y


class _Multi:
    geom_type = "MultiPoint"
    is_empty = False

    def __init__(self, pts):
        self.geoms = [_Point(*p) for p in pts]


class _Empty:
    geom_type = "LineString"
    is_empty = True
    coords = []


DEGENERATE = {"flag": False}


class LineString:
    def __init__(self, coords):
        pts = [np.asarray(c, dtype=np.float64) for c in coords]
        if len(pts) == 1:
            raise GEOSException("IllegalArgumentException: point array must contain 0 or >1 elements")
        self.pts = pts

    def intersection(self, other):
        out = []
        for i in range(len(self.pts) - 1):
            p, r = self.pts[i], self.pts[i + 1] - self.pts[i]
            for j in range(len(other.pts) - 1):
                q, s = other.pts[j], other.pts[j + 1] - other.pts[j]
                den = r[0] * s[1] - r[1] * s[0]
                qp = q - p
                if den == 0.0:
                    if qp[0] * r[1] - qp[1] * r[0] == 0.0 and (r @ r > 0 or s @ s > 0):
                        DEGENERATE["flag"] = True          # collinear: GEOS may return a LineString; scene dropped
                    continue
                t = (qp[0] * s[1] - qp[1] * s[0]) / den
                u = (qp[0] * r[1] - qp[1] * r[0]) / den
                eps = 1e-9
                if -eps <= t <= 1 + eps and -eps <= u <= 1 + eps:
                    if min(abs(t), abs(t - 1), abs(u), abs(u - 1)) < 1e-9:
                        DEGENERATE["flag"] = True          # through a vertex: predicate-dependent; scene dropped
                    out.append((float(p[0] + t * r[0]), float(p[1] + t * r[1])))
        uniq = sorted(set(out))
        if not uniq:
            return _Empty()
        return _Point(*uniq[0]) if len(uniq) == 1 else _Multi(uniq)


def _install_others():
    sh = types.ModuleType("shapely")
    sh.LineString = LineString
    err = types.ModuleType("shapely.errors")
    err.GEOSException = GEOSException
    sh.errors = err
    sys.modules["shapely"], sys.modules["shapely.errors"] = sh, err
    gym = types.ModuleType("gymnasium")
    gym.Env = object
    sys.modules["gymnasium"] = gym
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


class _Env:
    """What Agent.__init__ reads from the gym env (agents/base_agent.py:28-35; config/config.py:4-44)."""

    def __init__(self, vehicles_count):
        self.unwrapped = self
        self.config = {"simulation_frequency": 30, "policy_frequency": 10, "observation": {"vehicles_count": vehicles_count}}


def make_agent(module, vehicles_count, horizon=20):
    cfg = {"horizon": horizon, "render": False, "weight_speed": 1.0, "weight_control": 1.0, "weight_input_diff": 1.0,
           "speed_override": None}
    ag = module.PureMPC_Agent(_Env(vehicles_count), cfg)
    orig = ag.update_reference_states

    def recording(*a, **k):
        ref = orig(*a, **k)
        CAPTURE["ref_speed_column"] = np.array(ref[:, 2], dtype=np.float64)
        return ref
    ag.update_reference_states = recording
    return ag


def run_reference_step(ag, obs, ref_speed, U, rollout):
    """One `predict` of the reference with (X, U) injected.  X is rolled out from the state the oracle parses
    (so a parse mismatch shows up as a non-zero initial-condition residual in g)."""
    N = U.shape[0]
    X = rollout(obs, U)
    INJECT["x"], INJECT["u"] = X.T.copy(), U.T.copy()
    CAPTURE.clear()
    DEGENERATE["flag"] = False
    act = ag.predict(obs, return_numpy=True, weights_from_RL=None,
                     ref_speed=None if ref_speed is None else np.array([[ref_speed]], dtype=np.float32))
    n_obs = len(ag.agent_vehicles)
    flags = np.zeros(n_obs, np.uint8)
    cidx = -np.ones(n_obs, np.int32)
    have_lists = hasattr(ag, "agent_collide") and len(getattr(ag, "agent_collide", [])) == n_obs
    if have_lists:
        flags[:] = np.array(ag.agent_collide, dtype=bool)
    ci = getattr(ag, "conflict_index", None)
    if ci is not None and len(ci) == n_obs:
        cidx[:] = [-1 if c is None else int(c) for c in ci]
    return dict(X=X, action=np.asarray(act, dtype=np.float64), f=CAPTURE["f"], components=np.array(CAPTURE["components"]),
                g_max=float(np.max(np.abs(CAPTURE["g"]))), g0=CAPTURE["g"][:4].copy(), lbx=CAPTURE["lbx"], ubx=CAPTURE["ubx"],
                x0=CAPTURE["x0"], ref_v=CAPTURE["ref_speed_column"], ego_index=int(ag.ego_index),
                is_collide=bool(ag.is_collide), memory=int(ag.collision_memory), flags=flags, cidx=cidx,
                ego_state=np.array([ag.ego_vehicle.position[0], ag.ego_vehicle.position[1], ag.ego_vehicle.heading, ag.ego_vehicle.speed],
                                   dtype=np.float64),
                degenerate=bool(DEGENERATE["flag"]))


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs the reference tree at /root/reference")
    _install_casadi()
    _install_others()
    sys.path.insert(0, REF)
    import agents.pure_mpc as ref_mpc                      # the reference, unmodified
    import agents.archive.pure_mpc as ref_archive          # the only version whose objective has the obstacle-distance term
    import agents.pure_mpc_no_collision as ref_nocoll      # BASELINE config 2 names it; its objective omits the state cost (quirk Q3)
    import helpers
    from helpers import orc
    import mpc_rl_for_avs_b200 as pkg

    N, M = 20, 8
    V = M + 1
    rng = np.random.default_rng(20240611)

    def rollout(obs, U):
        p = orc.parse_obs(obs, V)
        return orc.rollout(np.array(p.ego, dtype=np.float64), U)

    # ---- single-step scenes, cold latch --------------------------------------------------------------------
    S = 192
    obs_all, rs_all, has_all = pkg.make_scenarios(S, M, seed=4321)
    obs_all, rs_all, has_all = obs_all.numpy(), rs_all.numpy().reshape(-1), has_all.numpy().reshape(-1)
    # some edge rows: ego beyond the path end (short polyline), absent vehicles, heading outside [-pi, pi]
    obs_all[0, 0, 1:3] = (-23.4, 14.3)
    obs_all[1, 5:, 0] = 0.0
    obs_all[2, 0, 5] += 2 * np.pi
    obs_all[3, 0, 3:5] = 0.0
    keep, rec = [], []
    for i in range(S):
        ag = make_agent(ref_mpc, V, N)
        U = np.stack([rng.uniform(-3, 3, N), rng.uniform(-0.3, 0.3, N)], axis=1)
        if i % 4 == 0:
            U[:] = 0.0                                      # the reference's own starting point
        r = run_reference_step(ag, obs_all[i], float(rs_all[i]) if has_all[i] else None, U, rollout)
        if r["degenerate"]:
            continue
        r["U"] = U
        # distance component of agents/archive/pure_mpc.py:189-196 at the same (X, U): parse + _solve of the archive agent
        arch = ref_archive.PureMPC_Agent(_Env(V), {"horizon": N, "render": False, "speed_override": None, "weight_state": 10.0,
                                                   "weight_control": 1.0, "weight_distance": 10.0, "weight_collision": 0.0,
                                                   "weight_input_diff": 1.0, "weight_final_state": 0.0})
        arch._parse_obs(obs_all[i])
        arch.is_collide = False
        CAPTURE.clear()
        arch._solve()
        r["distance_component"] = CAPTURE["components"][4]
        r["archive_f"] = CAPTURE["f"]
        # the shipped agent again with RL-set weights (v1 agents: weights_from_RL [[speed, control, input_diff]],
        # agents/pure_mpc.py:96-104) on a fresh agent object; is_collide still forces the speed weight to 100 (quirk Q9)
        wrl = rng.uniform(0.1, 5.0, size=(1, 3))
        agw = make_agent(ref_mpc, V, N)
        INJECT["x"], INJECT["u"] = r["X"].T.copy(), U.T.copy()
        CAPTURE.clear()
        agw.predict(obs_all[i], return_numpy=True, weights_from_RL=wrl, ref_speed=None)
        r["rl_weights"] = wrl[0].copy()
        r["rl_weights_f"] = CAPTURE["f"]
        INJECT["x"], INJECT["u"] = r["X"].T.copy(), U.T.copy()
        # agents/pure_mpc_no_collision.py at the same (X, U): its literal objective (control + input_diff only), with the
        # RL reference speed when the scene has one
        nc = ref_nocoll.PureMPC_Agent(_Env(V), {"horizon": N, "render": False, "weight_speed": 1.0, "weight_control": 1.0,
                                                "weight_input_diff": 1.0, "speed_override": None})
        CAPTURE.clear()
        nc.predict(obs_all[i], return_numpy=True, weights_from_RL=None,
                   ref_speed=np.array([[rs_all[i]]], dtype=np.float32) if has_all[i] else None)
        r["nocoll_f"] = CAPTURE["f"]
        r["nocoll_gmax"] = float(np.max(np.abs(CAPTURE["g"])))
        r["nocoll_ego_index"] = int(nc.ego_index)
        keep.append(i)
        rec.append(r)
    out = dict(obs=obs_all[keep], ref_speed=rs_all[keep], has_ref_speed=has_all[keep])
    for k in ("U", "X", "action", "f", "components", "g_max", "g0", "ref_v", "ego_index", "is_collide", "memory", "ego_state",
              "distance_component", "archive_f", "nocoll_f", "nocoll_gmax", "nocoll_ego_index",
              "rl_weights", "rl_weights_f"):
        out["ss_" + k] = np.array([r[k] for r in rec])
    out["ss_flags"] = np.array([np.pad(r["flags"], (0, M - len(r["flags"]))) for r in rec])
    out["ss_cidx"] = np.array([np.pad(r["cidx"], (0, M - len(r["cidx"])), constant_values=-1) for r in rec])
    out["lbx"], out["ubx"] = rec[0]["lbx"], rec[0]["ubx"]
    out["ss_x0"] = np.array([r["x0"] for r in rec])

    # ---- latch sequences: one agent object stepped through a moving scene ------------------------------------
    Q, T = 24, 16
    obs0, _, _ = pkg.make_scenarios(Q, M, seed=777)
    obs0 = obs0.numpy()
    path = pkg.reference_path(0.1)
    seq_obs = np.zeros((Q, T, V, 8), np.float32)
    seq = {k: [] for k in ("is_collide", "memory", "ego_index", "ref_v", "flags", "cidx")}
    ok = np.ones(Q, bool)
    for q in range(Q):
        ag = make_agent(ref_mpc, V, N)
        o = obs0[q].astype(np.float64)
        j0 = int(rng.integers(0, 30))
        rows = {k: [] for k in seq}
        for t in range(T):
            j = min(j0 + t, 84)                              # ego follows the path one point per step
            o[0, 1:3] = path[j, :2] + (0.05, -0.03)
            o[0, 5] = path[j, 3]
            sp = 9.0 if t < 8 else 4.0
            o[0, 3:5] = (sp * np.cos(o[0, 5]), sp * np.sin(o[0, 5]))
            o[0, 6:8] = (np.sin(o[0, 5]), np.cos(o[0, 5]))
            if t > 0:
                o[1:, 1:3] += 0.1 * o[1:, 3:5]              # others at constant velocity
            ob = o.astype(np.float32)
            seq_obs[q, t] = ob
            r = run_reference_step(ag, ob, None, np.zeros((N, 2)), rollout)
            ok[q] &= not r["degenerate"]
            for k in rows:
                rows[k].append(np.pad(r[k], (0, M - len(r[k])), constant_values=(-1 if k == "cidx" else 0)) if k in ("flags", "cidx") else r[k])
        for k in seq:
            seq[k].append(np.array(rows[k]))
    out["seq_obs"] = seq_obs[ok]
    for k in seq:
        out["seq_" + k] = np.array(seq[k])[ok]
    path_out = os.path.join(HERE, "golden_reference.npz")
    np.savez_compressed(path_out, **out)
    print("wrote", path_out, "single-step scenes", len(keep), "of", S, "| sequences", int(ok.sum()), "of", Q,
          "| collide frac", float(np.mean(out["ss_is_collide"])), "| max |g|", float(np.max(out["ss_g_max"])))


if __name__ == "__main__":
    main()
