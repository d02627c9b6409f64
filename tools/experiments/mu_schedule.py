"""Levenberg schedule sweep on the host build of the device code (2048 problems of the bench workload)."""
import sys, subprocess, numpy as np, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import helpers
d = dict(np.load("/tmp/batch2048.npz"))
B = d["ego_index"].shape[0]
def run(defs, tag):
    so = f"/tmp/libhostsim_{tag}.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", *defs, "-shared", "-fPIC", "-o", so, "/root/repo/tests/hostsim/hostsim.cpp"])
    lib = C.CDLL(so)
    cfg = helpers.hs_config(N=20, M=8, w_distance=10.0)
    act=np.zeros((B,2),np.float32); st=np.zeros(B,np.int32); it=np.zeros(B,np.int32); cost=np.zeros(B,np.float32); U=np.zeros((B,20,2),np.float32); fails=np.zeros(B,np.int32)
    b = helpers._as_struct(d); P = helpers._P
    f = lambda a, t: a.ctypes.data_as(P(t))
    lib.hs_solve(C.byref(cfg), helpers.REF.ctypes.data_as(P(C.c_double)), C.byref(b), B, 0, f(act,C.c_float), f(st,C.c_int), f(it,C.c_int), f(cost,C.c_float), f(U,C.c_float), f(fails,C.c_int))
    return dict(it=it, st=st, cost=cost, fails=fails, act=act)
base = run([], "base")
def report(tag, r):
    rel = (r["cost"] - base["cost"]) / (1 + np.abs(base["cost"]))
    print("%-14s mean_it %.2f accepted/problem %.2f rej %.3f conv %.3f cap %.3f  >30it %.3f | cost worse>1e-3: %.3f better>1e-3: %.3f | same u0: %.3f" % (
        tag, r["it"].mean(), (r["it"] - r["fails"]).mean(), r["fails"].sum() / r["it"].sum(), (r["st"] == 0).mean(), ((r["st"] & 1) != 0).mean(),
        (r["it"] > 30).mean(), (rel > 1e-3).mean(), (rel < -1e-3).mean(), (np.abs(r["act"] - base["act"]).max(1) < 1e-3).mean()))
report("base 0.1/30/3", base)
for dec, inc, mn in ((0.2, 30, 3), (0.3, 10, 3), (0.3, 30, 3), (0.1, 10, 3), (0.1, 30, 1), (0.1, 30, 10), (0.2, 20, 3), (0.5, 10, 3)):
    r = run([f"-DMPC_MU_DEC={dec}", f"-DMPC_MU_INC={inc}", f"-DMPC_MU_MIN={mn}"], f"{dec}_{inc}_{mn}")
    report(f"{dec}/{inc}/{mn}", r)
