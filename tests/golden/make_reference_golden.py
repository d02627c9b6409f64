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


class _Line:
    geom_type = "LineString"
    is_empty = False

    def __init__(self, coords):
        self.coords = list(coords)


class _MultiLine:
    geom_type = "MultiLineString"
    is_empty = False

    def __init__(self, lines):
        self.geoms = [_Line(c) for c in lines]


class _Collection:
    geom_type = "GeometryCollection"          # pure_mpc.py:615-633 has no branch for it
    is_empty = False

    def __init__(self, geoms):
        self.geoms = geoms


DEGENERATE = {"flag": False}


class LineString:
    """Stand-in for shapely.geometry.LineString: textbook parametric segment intersection.  Collinear overlaps are
    noded like an overlay would (every vertex of either line inside the overlap becomes a coordinate of the result) and
    chained along `self`; exact vertex touches are points; anything within 1e-9 of a vertex without being exact marks the
    scene as predicate-dependent (dropped from the fixture)."""

    def __init__(self, coords):
        pts = [np.asarray(c, dtype=np.float64) for c in coords]
        if len(pts) == 1:
            raise GEOSException("IllegalArgumentException: point array must contain 0 or >1 elements")
        self.pts = pts

    def intersection(self, other):
        out = []
        edges = []                                          # noded collinear sub-segments, as (a, b) vertex tuples along self
        for i in range(len(self.pts) - 1):
            p, r = self.pts[i], self.pts[i + 1] - self.pts[i]
            for j in range(len(other.pts) - 1):
                q, s = other.pts[j], other.pts[j + 1] - other.pts[j]
                den = r[0] * s[1] - r[1] * s[0]
                qp = q - p
                if den == 0.0:
                    if qp[0] * r[1] - qp[1] * r[0] == 0.0 and r @ r > 0 and s @ s > 0:
                        # collinear: overlap of [p, p+r] and [q, q+s] measured along r; the overlap's ends are two of the
                        # four vertices
                        four = [(0.0, tuple(self.pts[i])), (1.0, tuple(self.pts[i + 1])),
                                (float(qp @ r) / float(r @ r), tuple(other.pts[j])),
                                (float((qp + s) @ r) / float(r @ r), tuple(other.pts[j + 1]))]
                        lo = max(0.0, min(four[2][0], four[3][0]))
                        hi = min(1.0, max(four[2][0], four[3][0]))
                        if hi > lo:
                            ends = sorted([f for f in four if lo <= f[0] <= hi], key=lambda f: f[0])
                            edges.append((ends[0][1], ends[-1][1]))
                        elif hi == lo:
                            out.append([f[1] for f in four if f[0] == lo][0])
                    elif r @ r == 0 or s @ s == 0:
                        DEGENERATE["flag"] = True          # zero-length segment: scene dropped
                    continue
                t = (qp[0] * s[1] - qp[1] * s[0]) / den
                u = (qp[0] * r[1] - qp[1] * r[0]) / den
                eps = 1e-9
                if -eps <= t <= 1 + eps and -eps <= u <= 1 + eps:
                    if t in (0.0, 1.0) and 0.0 <= u <= 1.0:
                        out.append(tuple(self.pts[i] if t == 0.0 else self.pts[i + 1]))      # exact vertex touch
                    elif u in (0.0, 1.0) and 0.0 <= t <= 1.0:
                        out.append(tuple(other.pts[j] if u == 0.0 else other.pts[j + 1]))
                    elif min(abs(t), abs(t - 1), abs(u), abs(u - 1)) < 1e-9:
                        DEGENERATE["flag"] = True          # next to a vertex: predicate-dependent; scene dropped
                    else:
                        out.append((float(p[0] + t * r[0]), float(p[1] + t * r[1])))
        lines = []
        for a, b in edges:                                  # chain the noded edges in the order of `self`
            if lines and lines[-1][-1] == a:
                lines[-1].append(b)
            elif lines and a in lines[-1] and b in lines[-1]:
                pass
            else:
                lines.append([a, b])
        # the sub-segments of one overlap arrive once per (i, j) pair: restore every vertex in between
        merged = []
        for ln in lines:
            lo_, hi_ = ln[0], ln[-1]
            d = np.subtract(hi_, lo_)
            along = lambda v: float(np.subtract(v, lo_) @ d)  # noqa: E731
            verts = {tuple(v) for v in self.pts + other.pts
                     if (v[0] - lo_[0]) * d[1] - (v[1] - lo_[1]) * d[0] == 0.0 and 0.0 <= along(v) <= along(hi_)}
            merged.append(sorted(verts, key=along))
        on_line = lambda pt: any(pt in ln for ln in merged)  # noqa: E731
        uniq = sorted({pt for pt in out if not on_line(pt)})
        if merged and uniq:
            return _Collection([_Point(*u_) for u_ in uniq] + [_Line(c) for c in merged])
        if merged:
            return _Line(merged[0]) if len(merged) == 1 else _MultiLine(merged)
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
    # same-lane traffic: every fourth scene gets one or two vehicles exactly on the lane centre x = 2.0 (the path's own x)
    # heading -pi/2, ahead of or behind the ego -- the collinear / LineString branch of pure_mpc.py:615-633
    for i in range(4, S, 4):
        for m in (1, 2)[: 1 + (i // 4) % 2]:
            sp = float(rng.uniform(2.0, 9.0))
            y = float(np.float32(obs_all[i, 0, 2] + rng.uniform(-25.0, 25.0)))
            obs_all[i, m] = (1.0, 2.0, y, np.float32(sp * np.cos(-np.pi / 2)), -sp, -np.pi / 2, -1.0, np.float32(np.cos(-np.pi / 2)))
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
    for q in range(0, Q, 3):                                # a lead vehicle on the ego's own lane centre in every third sequence
        obs0[q, 1] = (1.0, 2.0, np.float32(20.0 + q), np.float32(5.0 * np.cos(-np.pi / 2)), -5.0, -np.pi / 2, -1.0, np.float32(np.cos(-np.pi / 2)))
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
