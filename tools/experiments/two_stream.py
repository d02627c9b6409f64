"""Does overlapping consecutive (independent) steps on several streams hide the tail of the persistent solve kernel?
One BatchedPureMPC handle per stream, steps issued round-robin; CUDA events around K steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mpc_rl_for_avs_b200 as pkg
B, M = 65536, 8
CFG = {"horizon": 20, "weight_speed": 1.0, "weight_control": 1.0, "weight_input_diff": 1.0}
dev = torch.device("cuda", 0)
batches = []
for s in range(8):
    obs, rs, has = pkg.make_scenarios(B, M, seed=1234 + 1000 * s)
    rsn = torch.where(has.reshape(-1, 1), rs, torch.full_like(rs, float("nan"))).reshape(-1).contiguous()
    batches.append((obs.contiguous().to(dev), rsn.to(dev)))
for n_starts in (1, 4):
    for depth in (1, 2, 3):
        agents = [pkg.BatchedPureMPC(CFG, vehicles_count=M + 1, max_batch=B, device=0, collision_check=True, weight_distance=10.0, n_starts=n_starts) for _ in range(depth)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        def step(i):
            k = i % depth
            with torch.cuda.stream(streams[k]):
                o, r = batches[i % 8]
                agents[k].reset()
                agents[k].predict_batch(o, ref_speed=r)
        for i in range(6): step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 24
        e0.record()
        for s in streams: s.wait_event(e0)
        for i in range(K): step(i)
        for s in streams: torch.cuda.current_stream(dev).wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"n_starts {n_starts} depth {depth}: {ms:.3f} ms/step  {B / ms * 1e3 / 1e6:.2f} M solves/s", flush=True)
        for a in agents: a.close()
