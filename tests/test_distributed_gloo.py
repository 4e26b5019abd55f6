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


def _update_worker(rank, world, port, q):
    """Two ranks with UNEQUAL shards (5 and 4 envs) run whole PPO updates: the bucketed gradient
    all-reduce is driven from the backward hooks, every rank performs the same number of
    optimisation steps, advantages are normalised with global moments, and the parameters stay
    bit-identical across ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root]
    from inversus_b200.policies import InversusCNNPolicy
    from inversus_b200.ppo_agent import DeviceRollout, PPOAgent
    from inversus_b200.sharding import shard_range
    torch.manual_seed(0)  # identical initial weights
    agent = PPOAgent(InversusCNNPolicy(), device="cpu", epochs=2, batch_size=16)
    _, n_local = shard_range(9, rank, world)  # 5 / 4 envs
    T = 7                                     # 35 vs 28 samples: 3 vs 2 minibatches of 16 without the fix
    g = torch.Generator().manual_seed(100 + rank)
    ro = DeviceRollout(T, n_local, "cpu", store="obs")
    steps = []
    for _ in range(2):  # two updates
        for t in range(T):
            obs = (torch.rand(n_local, 12, 10, 15, generator=g) > 0.7).float()
            extra = torch.rand(n_local, 4, generator=g)
            a, lp, v = agent.act(obs, extra)
            ro.store_pre((obs, extra), a, lp, v)
            ro.store_post(torch.randn(n_local, generator=g), (torch.rand(n_local, generator=g) < 0.1).to(torch.uint8))
        before = float(agent.optimizer.state_dict()["state"].get(0, {}).get("step", 0.0))  # a copy: the tensor is updated in place
        agent.update(ro, None, torch.zeros(n_local))
        steps.append(int(float(agent.optimizer.state_dict()["state"][0]["step"]) - before))
    flat = torch.cat([p.detach().reshape(-1) for p in agent.policy.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    assert agent._flat_grad is not None and len(agent._buckets) >= 2  # the overlapped path ran
    q.put((rank, same, steps, bool(torch.isfinite(flat).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_update_keeps_ranks_in_step_and_identical():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_update_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, same0, steps0, fin0), (_, same1, steps1, fin1) = res
    assert same0 and same1 and fin0 and fin1
    assert steps0 == steps1 == [6, 6]  # ceil(35 / 16) = 3 steps x 2 epochs on BOTH ranks
