"""The numpy drop-in (`inversus_b200.env_wrappers`) driven exactly the way the reference's trainer
drives `MultiEnvRunner` (inversus_rl/training.py:119-157 and :287-325), checked against the oracle
stepped with the same actions: observations, rewards, dones, info dicts, episode bookkeeping and
the per-env `envs[i].reset()` path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _trainer_style_rollout(runner, ref, steps, rs, selfplay=False, opponent_policy=None):
    """training.py:119-157 with the oracle mirrored step by step."""
    obs_grid, obs_extra = runner.reset()
    rg, re_ = ref.reset()
    assert obs_grid.dtype == np.float32 and obs_grid.shape == (runner.num_envs, 12, 10, 15)
    assert np.array_equal(obs_grid, rg) and np.array_equal(obs_extra, re_)
    episode_rewards, wins = [], []
    for _ in range(steps):
        actions = rs.randint(0, 13, size=runner.num_envs)
        a2 = None
        if selfplay:
            a2 = np.array([opponent_policy((ref.obs2[i], ref.extra2[i])) for i in range(runner.num_envs)])
        next_obs, rewards, dones, infos = runner.step(actions, opponent_policy)
        next_grid, next_extra = next_obs
        (rg, re_), rrew, rdone, rflags = ref.step(actions, a2, auto_reset=False)
        assert rewards.dtype == np.float32 and dones.dtype == bool and len(infos) == runner.num_envs
        assert np.array_equal(next_grid, rg) and np.array_equal(next_extra, re_)
        assert np.array_equal(rewards, rrew) and np.array_equal(dones, rdone)
        for i in range(runner.num_envs):
            info = infos[i]
            assert info["episode_steps"] == ref.episode_steps[i] and info["episode_return"] == ref.episode_return[i]
            assert info["win"] == bool(rflags[i] & 4) and info["lose"] == bool(rflags[i] & 8)
            assert info["landed_hit"] == bool(rflags[i] & 1) and info["got_hit"] == bool(rflags[i] & 2)
            if dones[i]:  # training.py:140-151
                episode_rewards.append(infos[i].get("episode_return", 0.0))
                wins.append(1 if infos[i].get("win", False) else 0)
                reset_obs = runner.envs[i].reset()
                next_grid[i] = reset_obs[0]          # the trainer mutates the returned arrays in place
                next_extra[i] = reset_obs[1]
                ref.reset_env(i)
        obs_grid, obs_extra = next_grid, next_extra
    return episode_rewards, wins


class _Ref:
    """OracleBatch plus single-env reset, so the trainer loop can be mirrored."""

    def __init__(self, n, mode, difficulty, max_steps, seed):
        from oracle import oracle as orc
        self.orc = orc
        self.b = orc.OracleBatch(n, mode, difficulty, max_steps, seed=seed)
        self.envs = self.b.envs

    def reset(self):
        out = self.b.reset()
        self.obs2, self.extra2 = self.b.obs2, self.b.extra2
        return out

    def reset_env(self, i):
        import ctypes as C
        L, b = self.orc.lib(), self.b
        L.orc_rl_reset(C.byref(b.envs[i]))
        L.orc_build_obs(C.byref(b.envs[i]), 0, b.obs1[i].ctypes.data, b.extra1[i].ctypes.data)
        if b.obs2 is not None:
            L.orc_build_obs(C.byref(b.envs[i]), 1, b.obs2[i].ctypes.data, b.extra2[i].ctypes.data)

    def step(self, a1, a2, auto_reset):
        out = self.b.step(a1, a2, auto_reset=auto_reset)
        self.episode_steps, self.episode_return = self.b.episode_steps, self.b.episode_return
        self.obs2, self.extra2 = self.b.obs2, self.b.extra2
        return out


def test_multienvrunner_in_the_reference_trainer_loop_vs_dummy():
    from inversus_b200 import MultiEnvRunner
    n = 4  # BASELINE.json configs[0]
    runner = MultiEnvRunner(n, opponent_type="dummy", max_episode_steps=60, difficulty="hard", seed=123)
    ref = _Ref(n, "dummy", "hard", 60, 123)
    assert runner.envs[0].env.width == 15 and runner.envs[0].env.height == 10
    ep, wins = _trainer_style_rollout(runner, ref, 400, np.random.RandomState(0))
    assert len(ep) > 10
    # MultiEnvRunner's own bookkeeping (env_wrappers.py:466-469, :513-519)
    assert sum(runner.episode_wins) == sum(wins)
    assert len(runner.episode_returns) == n and len(runner.episode_lengths) == n
    got, want = runner.sim.export_state(), ref.b.export_state()
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f]), f


def test_multienvrunner_selfplay_with_per_env_callback():
    """opponent_policy is called once per env with P2's view of the pre-step state
    (env_wrappers.py:311-314), exactly like the reference."""
    from inversus_b200 import MultiEnvRunner
    n = 6
    runner = MultiEnvRunner(n, opponent_type="selfplay", max_episode_steps=40, seed=9)
    ref = _Ref(n, "selfplay", "easy", 40, 9)
    seen = []

    def opponent_policy(obs):
        grid, extra = obs
        assert grid.shape == (12, 10, 15) and extra.shape == (4,)
        seen.append(1)
        # a deterministic function of P2's observation: shoot towards the enemy row/col, else move
        me = np.argwhere(grid[2] == 1)
        en = np.argwhere(grid[3] == 1)
        if len(me) == 0 or len(en) == 0:
            return 0
        (my, mx), (ey, ex) = me[0], en[0]
        if mx == ex:
            return 5 if ey < my else 7
        if my == ey:
            return 8 if ex < mx else 6
        return 2 if ex > mx else 4

    ref.reset()  # obs2 of the initial state for the first mirrored opponent call
    with pytest.raises(ValueError, match="opponent_policy required"):  # env_wrappers.py:309
        runner.reset()
        runner.step(np.zeros(n, np.int64))
    _trainer_style_rollout(runner, ref, 150, np.random.RandomState(1), selfplay=True, opponent_policy=opponent_policy)
    assert len(seen) >= 2 * 150 * n  # once per env per step on each side


def test_single_env_wrapper():
    """tests/test_rl_env_wrapper.py:53-110 of the reference, with the current 12-channel layout."""
    from inversus_b200 import SingleInversusRLEnv
    env = SingleInversusRLEnv(opponent_type="dummy", max_episode_steps=100, seed=4)
    grid, extra = env.reset()
    assert grid.shape == (12, env.env.height, env.env.width) and extra.shape == (4,)
    assert grid.dtype == np.float32 and extra.dtype == np.float32
    obs, reward, done, info = env.step(0)
    assert obs[0].shape == (12, 10, 15) and isinstance(reward, float) and isinstance(done, bool)
    assert set(info) == {"landed_hit", "got_hit", "win", "lose", "episode_steps", "episode_return"}
    assert info["episode_steps"] == 1 and env.step_count == 1
    for _ in range(120):
        obs, reward, done, info = env.step(1)
    assert done and info["episode_steps"] >= 100  # timeout flag stays up until the caller resets
    with pytest.raises(ValueError):
        env.step(13)


def test_build_observation_for_both_players():
    """`build_observation(runner.envs[i].env, PlayerId.P2)` (training.py:10 imports it): both
    perspectives of one env, equal to what the step wrote / to the oracle's P2 view."""
    from inversus_b200 import MultiEnvRunner, PlayerId, build_observation
    r = MultiEnvRunner(5, opponent_type="selfplay", max_episode_steps=50, seed=2)
    ref = _Ref(5, "selfplay", "easy", 50, 2)
    g, e = r.reset()
    ref.reset()
    rs = np.random.RandomState(0)
    for _ in range(12):
        a1, a2 = rs.randint(0, 13, 5), rs.randint(0, 13, 5)
        (g, e), _, _, _ = r.step(a1, opponent_actions=a2)
        ref.step(a1, a2, auto_reset=False)
    for i in range(5):
        g1, e1 = build_observation(r.envs[i].env, PlayerId.P1)
        g2, e2 = build_observation(r.envs[i].env, PlayerId.P2)
        assert np.array_equal(g1, g[i]) and np.array_equal(e1, e[i])
        assert np.array_equal(g2, ref.obs2[i]) and np.array_equal(e2, ref.extra2[i])
    with pytest.raises(ValueError):
        build_observation(r.envs[0].env, 3)


def test_reset_with_seed_reseeds_the_spawn_like_the_reference():
    """reference env_wrappers.py:272-276 / core.py:64-67: `reset(seed=s)` makes the new episode's
    spawn a function of s alone. Two runners with different seeds agree after `reset(seed=s)`,
    different s give different layouts, and the state equals the oracle's reset fed the same draws."""
    from inversus_b200 import MultiEnvRunner, SingleInversusRLEnv
    from inversus_b200.env_wrappers import _seeded_reset_draws
    from oracle import oracle as orc
    a = MultiEnvRunner(6, opponent_type="dummy", difficulty="hard", max_episode_steps=50, seed=11)
    b = MultiEnvRunner(3, opponent_type="dummy", difficulty="hard", max_episode_steps=50, seed=9999)
    a.reset()
    b.reset()
    rs = np.random.RandomState(0)
    for _ in range(7):
        a.step(rs.randint(0, 13, 6))
    before = a.sim.export_state()
    ga, ea = a.envs[4].reset(seed=77)
    gb, eb = b.envs[1].reset(seed=77)
    assert np.array_equal(ga, gb) and np.array_equal(ea, eb)
    after = a.sim.export_state()
    for i in (0, 1, 2, 3, 5):  # the other envs are untouched
        for f in after.dtype.names:
            assert np.array_equal(after[f][i], before[f][i]), (i, f)
    assert after["step_count"][4] == 0 and after["n_bullets"][4] == 0 and after["episode_return"][4] == 0.0
    layouts = {a.envs[0].reset(seed=s)[0].tobytes() for s in range(12)}
    assert len(layouts) > 6  # different seeds, different spawns
    # the oracle fed the same draws produces the same state
    ref = orc.OracleBatch(1, "dummy", "hard", 50, seed=11)
    table = np.zeros((1, 64), np.uint32)
    table[0, 16:60] = _seeded_reset_draws(77)
    ref.reset(table=table)
    want = ref.export_state()
    for f in ("tiles", "p1", "p2", "n_bullets", "step_count"):
        assert np.array_equal(after[f][4], want[f][0]), f
    # later plain resets go back to the runner's own stream
    g1, _ = a.envs[4].reset()
    g2, _ = a.envs[4].reset()
    assert g1.shape == (12, 10, 15) and not np.array_equal(g1, g2)
    env = SingleInversusRLEnv(opponent_type="dummy", max_episode_steps=20, seed=3)
    g, e = env.reset(seed=77)
    assert np.array_equal(g, ga) and np.array_equal(e, ea)
