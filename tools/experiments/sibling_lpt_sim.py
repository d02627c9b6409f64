"""Would ordering the portfolio's later starts by the iteration count of the problem's FIRST start shorten the launch?
List-scheduling simulation on per-start iteration counts of the bench workload (host build of the device code).

    python tools/experiments/sibling_lpt_sim.py [problems]
"""
import heapq, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import helpers, ipm_oracle as ipm
import mpc_rl_for_avs_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cache = f"/tmp/sibling_iters_{B}.npy"
if os.path.exists(cache):
    its = np.load(cache)
else:
    obs, rs, has = pkg.make_scenarios(B, 8, seed=1234)
    probs, _ = helpers.problems_from_obs(obs.numpy(), rs.numpy(), has.numpy(), w_distance=10.0, collision_check=True)
    d = helpers.batch_from_problems(probs, 8)
    lib = helpers.load_hostsim()
    its = []
    for st in range(4):
        U0 = np.stack([ipm.start_controls(st, 20)] * B).astype(np.float32)   # any four starts do for a scheduling study
        its.append(helpers.hostsim_solve_init(lib, d, helpers.hs_config(N=20, M=8, w_distance=10.0), U0=U0, n_starts=1)["iters"])
    its = np.stack(its)            # [4, B]
    np.save(cache, its)
print("mean iters per start", its.mean(1), "corr with start 0:", [round(float(np.corrcoef(its[0], its[k])[0, 1]), 3) for k in range(1, 4)])
long0 = its[0] >= 30
print("P(start k >= 30 | start0 >= 30) =", [round(float((its[k][long0] >= 30).mean()), 3) for k in range(1, 4)], " base rate", [round(float((its[k] >= 30).mean()), 3) for k in range(1, 4)])

lanes = int(round(37888 * B / 65536))


def makespan(order_items):
    """greedy list scheduling: each lane takes the next item when free; durations in trips"""
    h = [0.0] * lanes
    heapq.heapify(h)
    end = 0.0
    for (s, b) in order_items:
        t = heapq.heappop(h)
        t2 = t + its[s, b]
        end = max(end, t2)
        heapq.heappush(h, t2)
    return end


base = [(s, b) for s in range(4) for b in range(B)]
print("lanes", lanes, "total trips / lanes (lower bound)", its.sum() / lanes, "longest item", its.max())
print("start-major order (today)           makespan", makespan(base))
# oracle LPT over all items
allitems = sorted(base, key=lambda sb: -its[sb[0], sb[1]])
print("LPT with perfect knowledge          makespan", makespan(allitems))


def dynamic(priority):
    """event-driven: start-0 items in natural order first; a problem's other starts become available when its start 0
    finishes (that is when its iteration count is known) and are taken in order of `priority(iters of start 0)`"""
    free = [(0.0, l) for l in range(lanes)]
    heapq.heapify(free)
    posts = []                      # (time, problem) of finished start-0 items
    avail = []                      # (-priority, seq, start, problem)
    nxt0, done, end, seq = 0, 0, 0.0, 0
    total = 4 * B
    while done < total:
        t, l = heapq.heappop(free)
        while posts and posts[0][0] <= t:
            _, b = heapq.heappop(posts)
            for s_ in range(1, 4):
                heapq.heappush(avail, (-priority(its[0, b]), seq, s_, b)); seq += 1
        if nxt0 < B:
            s_, b = 0, nxt0; nxt0 += 1
        elif avail:
            _, _, s_, b = heapq.heappop(avail)
        else:                       # nothing available yet: wait for the next posting
            heapq.heappush(free, (posts[0][0], l))
            continue
        t2 = t + its[s_, b]
        if s_ == 0:
            heapq.heappush(posts, (t2, b))
        end = max(end, t2); done += 1
        heapq.heappush(free, (t2, l))
    return end


print("dynamic, FIFO of postings            makespan", dynamic(lambda i0: 0))
print("dynamic, sibling LPT (by start-0 its) makespan", dynamic(lambda i0: float(i0)))
for T in (20, 30, 40):
    print(f"dynamic, two classes (start0 >= {T})   makespan", dynamic(lambda i0, T=T: float(i0 >= T)))
