"""world_size-2 CPU test (gloo) of the multi-GPU layout: each rank owns a contiguous shard of the
global env ids, steps it with NO communication, and only the four rollout statistics are reduced.
The oracle stands in as the checker for what each shard must contain (RNG keyed by global id)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    from inversus_b200.sharding import dist_env, max_over_ranks, reduce_rollout_stats, shard_range
    from oracle import oracle as orc
    assert dist_env()[0] == rank and dist_env()[2] == world
    first, count = shard_range(total, rank, world)
    acts = np.random.RandomState(1).randint(0, 13, size=(T, total)).astype(np.int8)
    b = orc.OracleBatch(count, "dummy", "hard", 50, seed=77, env_id_base=first)
    b.reset()
    episodes = wins = 0
    ret_sum = len_sum = 0.0
    for t in range(T):
        _, _, done, flags = b.step(acts[t, first:first + count], auto_reset=True, want_obs=False)
        episodes += int(done.sum())
        wins += int(((flags & 4) != 0).sum())
        ret_sum += float(b.episode_return[done].sum())
        len_sum += float(b.episode_steps[done].sum())
    stats = reduce_rollout_stats(episodes, wins, ret_sum, len_sum)
    slowest = max_over_ranks(float(rank + 1))
    state = b.export_state()
    q.put((rank, first, count, stats, slowest, state["p1"].copy(), state["tiles"].copy(), (episodes, wins, ret_sum, len_sum)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    total, T, world = 101, 120, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    from oracle import oracle as orc
    acts = np.random.RandomState(1).randint(0, 13, size=(T, total)).astype(np.int8)
    b = orc.OracleBatch(total, "dummy", "hard", 50, seed=77)
    b.reset()
    tot = np.zeros(4)
    for t in range(T):
        _, _, done, flags = b.step(acts[t], auto_reset=True, want_obs=False)
        tot += (done.sum(), ((flags & 4) != 0).sum(), b.episode_return[done].sum(), b.episode_steps[done].sum())
    whole = b.export_state()
    assert res[0][1] == 0 and res[0][2] + res[1][2] == total and res[1][1] == res[0][2]
    for rank, first, count, stats, slowest, p1, tiles, local in res:
        assert np.array_equal(whole["p1"][first:first + count], p1)      # shard == slice of the whole
        assert np.array_equal(whole["tiles"][first:first + count], tiles)
        np.testing.assert_allclose(stats, tot, rtol=1e-12)              # every rank sees the global sums
        assert slowest == float(world)                                   # max-over-ranks timing helper
    assert sum(r[7][0] for r in res) == tot[0]


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root]
    from inversus_b200.policies import InversusCNNPolicy
    from inversus_b200.ppo_agent import PPOAgent
    torch.manual_seed(0)
    agent = PPOAgent(InversusCNNPolicy(), device="cpu")
    for i, p in enumerate(agent.policy.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    agent._sync_grads()  # the trainer's gradient exchange: one flat all-reduce, then the mean
    ok = all(torch.allclose(p.grad, torch.full_like(p, (world + 1) / 2 * (i + 1)))
             for i, p in enumerate(agent.policy.parameters()))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_all_reduce_averages_over_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
