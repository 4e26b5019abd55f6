"""GPU checks of the policy's packed-state encoder (csrc/encoder_kernels.cu): block 1 of the policy
-- conv1 + bias + LayerNorm + ReLU, inversus_rl/policies.py:27-31,94 -- evaluated straight from the
80-byte env states must equal the same block applied by PyTorch (fp32) to the observation that
build_observation (env_wrappers.py:173-245) yields for those states; and the simulator's
observation-free mode must advance exactly like the observation-writing one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _states(n=700, steps=37, seed=5, mode="dummy"):
    """A sim stepped into a varied population: bullets in flight, dead players, fresh resets."""
    import torch
    from inversus_b200 import BatchedInversus
    sim = BatchedInversus(n, mode, "hard", 30, seed=seed, obs_dtype="f32", auto_reset=False, p2_view=True)
    sim.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    for _ in range(steps):
        a1 = torch.randint(0, 13, (n,), device="cuda", generator=g)
        a1 = torch.where(torch.rand(n, device="cuda", generator=g) < 0.5, torch.randint(5, 13, (n,), device="cuda", generator=g), a1)
        a2 = torch.randint(0, 13, (n,), device="cuda", generator=g) if mode == "selfplay" else None
        sim.step(a1.to(torch.int8), None if a2 is None else a2.to(torch.int8))
    return sim


def _block1_reference(obs, w1, b1, gamma, beta, eps=1e-5):
    import torch
    import torch.nn.functional as F
    z = F.conv2d(obs.float(), w1, b1, padding=1)
    return F.relu(F.layer_norm(z, (32, 10, 15), gamma, beta, eps))


def _params(seed=0):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    w1 = (torch.randn(32, 12, 3, 3, device="cuda", generator=g) * 0.2).requires_grad_(True)
    b1 = (torch.randn(32, device="cuda", generator=g) * 0.1).requires_grad_(True)
    gamma = (1.0 + 0.3 * torch.randn(32, 10, 15, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(32, 10, 15, device="cuda", generator=g)).requires_grad_(True)
    return w1, b1, gamma, beta


@pytest.mark.parametrize("mode", ["dummy", "selfplay"])
def test_encoder_forward_matches_block1_on_the_observation(mode):
    import torch
    from inversus_b200.fused_ops import PackedStates, encode_layer1
    torch.backends.cudnn.allow_tf32 = False
    sim = _states(mode=mode)
    w1, b1, gamma, beta = _params()
    snap = sim.snapshot()
    for view in (0, 1):
        obs, extra = sim.obs_from_packed(snap, view=view, obs_dtype="f32")
        with torch.no_grad():
            want = _block1_reference(obs, w1, b1, gamma, beta).permute(0, 2, 3, 1).reshape(sim.num_envs, -1)
            got, got_extra = encode_layer1(PackedStates(snap, view), w1, b1, gamma.permute(1, 2, 0).reshape(-1),
                                           beta.permute(1, 2, 0).reshape(-1))
        assert torch.equal(got_extra, extra)                          # exact: same table, same bits
        err = (got.float() - want).abs()
        # tolerance: the pre-LayerNorm map is kept in bf16 (8 mantissa bits, like the library path's
        # conv output) and so is the result: two roundings of 2^-9 relative, amplified by rstd*gamma
        # (up to ~3 here) on values up to ~6 => 6e-2 absolute worst case, 2e-3 on average
        assert err.max() < 6e-2, err.max()
        assert err.mean() < 2e-3
        assert ((got > 0) == (want > 0)).float().mean() > 0.995      # same ReLU pattern up to rounding at 0


def test_encoder_counts_coinciding_bullets_once_like_the_observation():
    import torch
    from inversus_b200 import BatchedInversus
    from inversus_b200.fused_ops import PackedStates, encode_layer1
    sim = BatchedInversus(4, "dummy", "hard", 30, seed=1, obs_dtype="f32", auto_reset=False, p2_view=True)
    sim.reset()
    st = sim.export_state()
    for i in range(4):
        st["n_bullets"][i] = 5
        st["bullets"][i, :5] = [[3, 4, 1, 0], [3, 4, 1, 0], [3, 4, 1, 1], [0, 0, 2, 1], [14, 9, 0, 0]]
    st["p1"][1, 4] = 0  # a dead P1
    sim.import_state(st)
    w1, b1, gamma, beta = _params(3)
    snap = sim.snapshot()
    for view in (0, 1):
        obs, _ = sim.obs_from_packed(snap, view=view, obs_dtype="f32")
        with torch.no_grad():
            want = _block1_reference(obs, w1, b1, gamma, beta).permute(0, 2, 3, 1).reshape(4, -1)
            got, _ = encode_layer1(PackedStates(snap, view), w1, b1, gamma.permute(1, 2, 0).reshape(-1),
                                   beta.permute(1, 2, 0).reshape(-1))
        assert (got.float() - want).abs().max() < 6e-2


def test_encoder_backward_matches_autograd_of_block1():
    import torch
    from inversus_b200.fused_ops import PackedStates, encode_layer1
    torch.backends.cudnn.allow_tf32 = False
    sim = _states(n=1500, steps=23, seed=9)
    snap = sim.snapshot()
    obs, _ = sim.obs_from_packed(snap, view=0, obs_dtype="f32")
    w1, b1, gamma, beta = _params(1)
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    dy = torch.randn(sim.num_envs, 4800, device="cuda", generator=g).to(torch.bfloat16)
    want = _block1_reference(obs, w1, b1, gamma, beta).permute(0, 2, 3, 1).reshape(sim.num_envs, -1)
    want.backward(dy.float())
    ref = [p.grad.clone() for p in (w1, b1, gamma, beta)]
    for p in (w1, b1, gamma, beta):
        p.grad = None
    got, _ = encode_layer1(PackedStates(snap, 0), w1, b1, gamma.permute(1, 2, 0).reshape(-1), beta.permute(1, 2, 0).reshape(-1))
    got.backward(dy)
    for name, p, r in zip(("w1", "b1", "gamma", "beta"), (w1, b1, gamma, beta), ref):
        a, b = p.grad.flatten().double(), r.flatten().double()
        cos = torch.dot(a, b) / (a.norm() * b.norm())
        rel = (a - b).norm() / b.norm()
        # bf16 feature map + bf16 dz in the kernel (as in the library path) vs an fp32 reference:
        # 2^-9 relative noise per element => ~1.5 % in norm on the 3456-element weight gradient
        assert cos > 0.9995 and rel < 3e-2, (name, float(cos), float(rel))
    # deterministic: same inputs, same bits
    first = [p.grad.clone() for p in (w1, b1, gamma, beta)]
    for p in (w1, b1, gamma, beta):
        p.grad = None
    got2, _ = encode_layer1(PackedStates(snap, 0), w1, b1, gamma.permute(1, 2, 0).reshape(-1), beta.permute(1, 2, 0).reshape(-1))
    got2.backward(dy)
    assert all(torch.equal(p.grad, f) for p, f in zip((w1, b1, gamma, beta), first))


def test_policy_on_packed_states_tracks_the_observation_path():
    import torch
    from inversus_b200.fused_ops import PackedStates
    from inversus_b200.policies import InversusCNNPolicy
    torch.manual_seed(0)
    m = InversusCNNPolicy().cuda()
    sim = _states(n=600, steps=19, seed=2)
    snap = sim.snapshot()
    obs, extra = sim.obs_from_packed(snap, view=0, obs_dtype="f32")
    with torch.no_grad():
        a = m(obs, extra)
        b = m.infer(PackedStates(snap, 0), None)
        c = m.infer(PackedStates(snap, 0).chunk(100, 300), None)
    assert (a[0] - b[0]).abs().max() < 3e-2 and (a[1] - b[1]).abs().max() < 3e-2
    # a slice of the batch goes through the same kernels (the library may pick another GEMM tiling
    # for the smaller batch, so equality holds to bf16 rounding, not bit for bit)
    assert (b[0][100:300] - c[0]).abs().max() < 1e-2 and (b[1][100:300] - c[1]).abs().max() < 1e-2
    # gradients through the packed path reach every parameter and agree with the observation path
    lo, va = m.forward_bf16(PackedStates(snap, 0), None)
    (lo.square().mean() + va.square().mean()).backward()
    gp = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad()
    lo, va = m.forward_bf16(obs, extra)
    (lo.square().mean() + va.square().mean()).backward()
    for k, p in m.named_parameters():
        a_, b_ = gp[k].flatten().double(), p.grad.flatten().double()
        assert torch.dot(a_, b_) / (a_.norm() * b_.norm() + 1e-30) > 0.99, k


def test_observation_free_step_advances_exactly_like_the_f32_one():
    import torch
    from inversus_b200 import BatchedInversus
    n = 3000
    for mode in ("dummy", "selfplay"):
        a = BatchedInversus(n, mode, "hard", 25, seed=7, obs_dtype="f32", auto_reset=True)
        b = BatchedInversus(n, mode, "hard", 25, seed=7, obs_dtype="none", auto_reset=True)
        assert b.obs is None and b.extra is not None
        a.reset()
        b.reset()
        g = torch.Generator(device="cuda")
        g.manual_seed(1)
        for _ in range(60):
            a1 = torch.randint(0, 13, (n,), device="cuda", generator=g).to(torch.int8)
            a2 = torch.randint(0, 13, (n,), device="cuda", generator=g).to(torch.int8) if mode == "selfplay" else None
            a.step(a1, a2)
            b.step(a1, a2)
            for k in ("extra", "reward", "done", "info", "episode_steps", "episode_return", "packed_state"):
                assert torch.equal(getattr(a, k), getattr(b, k)), k
            if mode == "selfplay":
                assert torch.equal(a.extra_p2, b.extra_p2)
        assert a.poll_status() == 0 and b.poll_status() == 0
        with pytest.raises(ValueError):
            b.step_host(np.zeros(n, np.int8), np.zeros(n, np.int8) if mode == "selfplay" else None,
                        {"obs": np.zeros((n, 12, 10, 15), np.float32)})
