"""N > 1 host logic on CPU: env-index sharding and the all-gather of actions, world_size 2, gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_ranges_cover_the_batch():
    from mpc_rl_for_avs_b200.sharding import shard_range
    for total in (0, 1, 7, 65536, 1 << 20):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpc_rl_for_avs_b200.sharding import shard_range, all_gather_actions
    lo, hi = shard_range(total, rank, world)
    # "actions" of this shard: a known function of the global env index
    idx = torch.arange(lo, hi, dtype=torch.float32)
    local = torch.stack([idx, -2 * idx], dim=1)
    full = all_gather_actions(local, total)
    # in-place form: the producer writes into its slice of the persistent gathered buffer
    from mpc_rl_for_avs_b200.sharding import ActionGather
    g = ActionGather(total, "cpu")
    assert g.local.shape[0] == hi - lo and g.local.data_ptr() == g.buffer[rank * g.n_max].data_ptr()
    for rep in range(2):                                  # the buffer is reused step after step
        g.local.copy_(local + rep)
        assert torch.equal(g.gather(), full + rep)
    # pipelined form (two buffers, the gather of step i overlaps step i + 1 on a GPU; synchronous on CPU)
    from mpc_rl_for_avs_b200.sharding import PipelinedActionGather
    pipe = PipelinedActionGather(total, "cpu")
    for step in range(5):
        gg = pipe.acquire(step)
        assert gg is pipe.slots[step % 2]
        gg.local.copy_(local + 10 * step)
        pipe.issue(step)
        if step > 0:
            assert torch.equal(pipe.result(step - 1), full + 10 * (step - 1))      # the previous step's result is still intact
    pipe.drain()
    assert torch.equal(pipe.result(4), full + 40)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_all_gather_actions_world2_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(total, dtype=np.float32)
    exp = np.stack([idx, -2 * idx], axis=1)
    for r in range(2):
        assert np.array_equal(got[r], exp)


class _StubMPC:
    def predict_batch(self, obs, ref_speed=None, weights=None, reset_mask=None):
        a = torch.zeros(obs.shape[0], 2)
        a[:, 0] = 0.2
        return a


def _rl_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpc_rl_for_avs_b200.rl import A2CMPC, PPOMPC, BatchedIntersectionEnv
    out = {}
    for name, Algo, kw in (("a2c", A2CMPC, {}), ("ppo", PPOMPC, {"n_epochs": 2, "batch_size": 16})):
        env = BatchedIntersectionEnv(8, 9, device="cpu", seed=100 + rank)          # each rank owns its own environments
        algo = Algo(env, _StubMPC(), n_steps=4, seed=rank, **kw)                   # different seeds: rank 0's weights are broadcast
        w0 = torch.cat([p.detach().flatten() for p in algo.policy.parameters()]).clone()
        algo.train_step()
        w1 = torch.cat([p.detach().flatten() for p in algo.policy.parameters()])
        out[name] = (w0.numpy(), w1.numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_rl_update_world2_gloo():
    """N2 over several ranks: environments are sharded, the policy is replicated (broadcast at start, gradients
    all-reduced per step), so all ranks hold identical weights before and after an update on different rollouts."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 77
    procs = [ctx.Process(target=_rl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for name in ("a2c", "ppo"):
        (a0, a1), (b0, b1) = got[0][name], got[1][name]
        assert np.array_equal(a0, b0)                          # broadcast of rank 0's initial weights
        assert np.allclose(a1, b1, rtol=0, atol=1e-7) and not np.array_equal(a0, a1)
