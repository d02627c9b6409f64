import sys, os, subprocess, numpy as np, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import helpers
import mpc_rl_for_avs_b200 as pkg
B = 2048
obs, rs, has = pkg.make_scenarios(B, 8, seed=1234)
it60 = np.load("/tmp/it_60.npy"); st60 = np.load("/tmp/st_60.npy")
hard = np.where(it60 >= 30)[0]
print("hard", len(hard), "status counts", {int(s): int((st60[hard] == s).sum()) for s in np.unique(st60[hard])})
sel = hard[: int(sys.argv[1]) if len(sys.argv) > 1 else 3]
probs, _ = helpers.problems_from_obs(obs.numpy()[sel], rs.numpy()[sel], has.numpy()[sel], w_distance=10.0, collision_check=True)
d = helpers.batch_from_problems(probs, 8)
so = "/tmp/libhostsim_trace.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMPC_TRACE", "-include", "cstdio", "-shared", "-fPIC", "-o", so, "/root/repo/tests/hostsim/hostsim.cpp"])
lib = C.CDLL(so)
cfg = helpers.hs_config(N=20, M=8, w_distance=10.0)
for j in range(len(sel)):
    dj = {k: np.ascontiguousarray(v[..., j:j+1]) if v.ndim > 1 else np.ascontiguousarray(v[j:j+1]) for k, v in d.items()}
    print("=== problem", sel[j], "is_collide", dj["is_collide"], "v0", dj["s0"][3], "vr", dj["vr_a"], dj["vr_slope"], dj["vr_b"], dj["vr_n"]); sys.stdout.flush()
    helpers.hostsim_solve(lib, dj, cfg)
