"""Fused policy step (b2d_policy_act, csrc/policy_kernels.cuh) against the oracle restatement of the
reference's per-step policy work (oracle/policy_oracle.py: models.Default.forward_eval + Normal
sample_logits + reward clamp + action clip; pinned against torch's own modules in test_policy_cpu.py).

Tolerances: precision="fp32" (float32 FMAs vs the float64 oracle): outputs of the MLP (mean, value) 2e-5
absolute for O(1) activations.  precision="tf32" (TF32 tensor-core GEMMs, the reference's own GPU setting,
pufferl.py:55): 2e-4 against the oracle with its GEMM operands rounded to TF32 the same way (the residual
is float32 accumulation order plus hidden units whose TF32 rounding flips on a 1e-7 GELU difference), and
5e-3 against the unrounded float64 oracle.  The noise itself 2e-6; log-prob 2e-4 absolute (it sums four
z^2/2 terms up to ~15); copies, clamps, clips and terminals are exact in both."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_policy_cpu import _random_policy, _weights  # noqa: E402


def _setup(rows, obs_dim=29, hidden=128, seed=0, **kw):
    from drone_b200.policy import FusedPolicyStep
    p = _random_policy(obs_dim, hidden, seed).cuda()
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    obs = torch.randn((rows, obs_dim), device="cuda", generator=g) * 1.2
    rew = torch.randn(rows, device="cuda", generator=g) * 2.0
    term = (torch.rand(rows, device="cuda", generator=g) < 0.1).to(torch.uint8)
    env_act = torch.full((rows, 4), 7.0, device="cuda")
    return p, obs, rew, term, env_act, FusedPolicyStep(p, obs, rew, term, env_act, **kw)


def _row(rows, obs_dim, k=1, K=3):
    """experience tensors [K, rows, ...]; returns the k-th row views (k=1: not 16-byte aligned for odd shapes)"""
    z = lambda *s: torch.full((K,) + s, -9.0, device="cuda")  # noqa: E731
    full = dict(observations=z(rows, obs_dim), actions=z(rows, 4), logprobs=z(rows), rewards=z(rows), terminals=z(rows), values=z(rows))
    return full, {n: t[k] for n, t in full.items()}


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("rows,obs_dim,hidden", [(1, 29, 128), (63, 29, 128), (65, 29, 8), (4133, 29, 128), (4133, 41, 128),
                                                 (20000, 41, 64), (3000, 29, 256), (100003, 29, 128)])
def test_fused_policy_step_matches_the_oracle(rows, obs_dim, hidden, precision):
    from oracle import policy_oracle as pol
    p, obs, rew, term, env_act, fused = _setup(rows, obs_dim, hidden, noise_seed=(5 << 32) | 77, row_id_base=1000,
                                               precision=precision)
    tol = 2e-5 if precision == "fp32" else 2e-4
    full, row = _row(rows, obs_dim)
    fused.seek(3)
    fused.act(**row)
    torch.cuda.synchronize()
    assert fused.calls == 4
    ref = pol.policy_act(_weights(p), obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), call=3, seed=(5 << 32) | 77,
                         row_id_base=1000, tf32_gemm=precision == "tf32")
    got = {n: t.cpu().numpy() for n, t in row.items()}
    assert np.array_equal(got["observations"].view(np.uint32), obs.cpu().numpy().view(np.uint32))
    assert np.array_equal(got["rewards"], np.clip(rew.cpu().numpy(), -1, 1))
    assert np.array_equal(got["terminals"], term.cpu().numpy().astype(np.float32))
    assert np.abs(got["values"] - ref["values"]).max() < tol * max(1.0, np.abs(ref["values"]).max())
    std = np.exp(_weights(p)["decoder_logstd"].astype(np.float64))
    z = (got["actions"].astype(np.float64) - ref["mean"]) / std
    assert np.abs(z - ref["noise"]).max() < 2.5 * tol  # mean error / std dominates; the noise alone is checked below
    assert np.abs(got["actions"] - ref["actions"]).max() < tol * max(1.0, np.abs(ref["actions"]).max())
    if precision == "tf32":  # and the TF32 result stays within TF32 distance of the unrounded model
        exact = pol.policy_act(_weights(p), obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), call=3, seed=(5 << 32) | 77,
                               row_id_base=1000)
        assert np.abs(got["values"] - exact["values"]).max() < 5e-3 * max(1.0, np.abs(exact["values"]).max())
        assert np.abs(got["actions"] - exact["actions"]).max() < 5e-3 * max(1.0, np.abs(exact["actions"]).max())
    assert np.abs(got["logprobs"] - ref["logprobs"]).max() < 2e-4
    assert np.array_equal(env_act.cpu().numpy(), np.clip(got["actions"], -1, 1))
    # rows of the experience tensors that were not addressed stay untouched
    for n, t in full.items():
        assert float(t[0].min()) == -9.0 and float(t[2].max()) == -9.0, n


def test_deterministic_mode_and_optional_stores():
    from oracle import policy_oracle as pol
    rows = 5000
    p, obs, rew, term, env_act, fused = _setup(rows, deterministic=True, precision="fp32")
    _, row = _row(rows, 29, k=0)
    fused.act(actions=row["actions"], logprobs=row["logprobs"])  # the other stores are skipped
    torch.cuda.synchronize()
    ref = pol.policy_act(_weights(p), obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), 0, 0, deterministic=True)
    assert np.abs(row["actions"].cpu().numpy() - ref["mean"]).max() < 2e-5 * max(1.0, np.abs(ref["mean"]).max())
    assert np.abs(row["logprobs"].cpu().numpy() - ref["logprobs"]).max() < 1e-5
    assert float(row["values"].min()) == -9.0 and float(row["observations"].max()) == -9.0
    assert np.array_equal(env_act.cpu().numpy(), np.clip(row["actions"].cpu().numpy(), -1, 1))


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_noise_stream_bits_calls_and_graph_replay(precision):
    """z recovered with zero weights (mean = 0, std = 1): the device Box-Muller against the float64 one;
    successive calls and CUDA-graph replays advance the call number; seek() rewinds it."""
    from oracle import policy_oracle as pol
    rows = 70001
    p, obs, rew, term, env_act, fused = _setup(rows, noise_seed=123456789, precision=precision)
    with torch.no_grad():
        p.decoder_mean.weight.zero_(), p.decoder_mean.bias.zero_(), p.decoder_logstd.zero_()
    acts = torch.zeros((4, rows, 4), device="cuda")
    lps = torch.zeros((4, rows), device="cuda")
    fused.act(actions=acts[0], logprobs=lps[0])
    fused.act(actions=acts[1], logprobs=lps[1])
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fused.act(actions=acts[2], logprobs=lps[2])  # call 2 (warm-up outside capture)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        fused.act(actions=acts[2], logprobs=lps[2])
    g.replay()  # call 3
    torch.cuda.synchronize()
    third = acts[2].clone()
    g.replay()  # call 4
    torch.cuda.synchronize()
    assert fused.calls == 5
    for call, got in ((0, acts[0]), (1, acts[1]), (3, third), (4, acts[2])):
        z = pol.noise(rows, call, 123456789)
        assert np.abs(got.cpu().numpy() - z).max() < 2e-6 * max(1.0, np.abs(z).max()), call
    z = acts[0].cpu().numpy().astype(np.float64)
    assert np.abs(lps[0].cpu().numpy() - (-0.5 * (z * z).sum(1) - 4 * 0.9189385332046727)).max() < 1e-5
    fused.seek(1)
    fused.act(actions=acts[3])
    torch.cuda.synchronize()
    assert torch.equal(acts[3], acts[1])


def test_gelu_accuracy_over_the_whole_range():
    """value = GELU(x) through a one-hot encoder row: |error| <= 3e-7 * max(1, |x|): 2.5 float32 ulps of 1 (the
    approximation itself is within 8e-8 of GELU in exact arithmetic; the rest is float32 rounding)."""
    from math import erf, sqrt
    rows = 200001
    p, obs, rew, term, env_act, fused = _setup(rows, precision="fp32")
    with torch.no_grad():
        for t in p.parameters():
            t.zero_()
        p.encoder[0].weight[5, 2] = 1.0
        p.value.weight[0, 5] = 1.0
        obs.zero_()
        obs[:, 2] = torch.linspace(-9.0, 9.0, rows, device="cuda")
    vals = torch.zeros(rows, device="cuda")
    fused.act(values=vals)
    torch.cuda.synchronize()
    x = obs[:, 2].cpu().numpy().astype(np.float64)
    exact = np.array([0.5 * v * (1.0 + erf(v / sqrt(2.0))) for v in x])
    err = np.abs(vals.cpu().numpy() - exact)
    assert (err <= 3e-7 * np.maximum(1.0, np.abs(x))).all(), err.max()
    tg = torch.nn.functional.gelu(obs[:, 2]).cpu().numpy()  # torch's own float32 GELU is not closer
    assert err.max() <= 4 * np.abs(tg - exact).max() + 2e-7


def test_argument_errors_raise_like_the_other_entry_points():
    from drone_b200.policy import FusedPolicyStep
    from drone_b200.rollout import DronePolicy
    rows = 100
    p, obs, rew, term, env_act, fused = _setup(rows)
    with pytest.raises(ValueError):
        fused.act(values=torch.zeros(rows + 1, device="cuda"))
    with pytest.raises(ValueError):
        fused.act(actions=torch.zeros((rows, 4), device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError):
        FusedPolicyStep(DronePolicy(obs_dim=30).cuda(), torch.zeros((rows, 30), device="cuda"), rew, term, env_act).act()
    with pytest.raises(ValueError):
        FusedPolicyStep(p, obs, rew, term, env_act, precision="fp16")
    with pytest.raises(ValueError):
        FusedPolicyStep(DronePolicy(hidden_size=100).cuda(), obs, rew, term, env_act).act()


def test_fused_rollout_equals_torch_rollout_on_the_first_steps():
    """DeviceRollout(policy_impl='fused') against policy_impl='torch' (deterministic policy, strict env):
    the stored experience agrees within the policy tolerance while trajectories have not diverged."""
    from drone_b200.rollout import DeviceRollout
    from drone_b200.vec import RaceVec
    n, K = 5000, 4
    p = _random_policy().cuda()
    outs = []
    for impl in ("torch", "fused"):
        vec = RaceVec(n, seed=4, math="strict")
        vec.reset(4)
        ro = DeviceRollout(vec, p, horizon=K, use_graph=False, deterministic=True, policy_impl=impl, precision="fp32")
        ro.collect()
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in ro.segments().items()})
        vec.close()
    a, b = outs
    assert torch.equal(a["observations"][:, 0], b["observations"][:, 0])
    for k in ("actions", "values", "logprobs", "rewards", "observations"):
        assert torch.allclose(a[k], b[k], rtol=1e-3, atol=2e-4), k
    assert (a["terminals"] != b["terminals"]).float().mean() < 1e-3


def test_swarm_rollout_with_the_fused_policy_in_a_graph():
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import SwarmVec
    torch.manual_seed(0)
    vec = SwarmVec(256, num_drones=16, max_rings=5, seed=1)
    vec.reset(1)
    p = DronePolicy(obs_dim=41).cuda()
    ro = DeviceRollout(vec, p, horizon=8, use_graph=True)
    assert ro.policy_impl == "fused"
    ro.collect()
    first = ro.actions.clone()
    ro.collect()
    torch.cuda.synchronize()
    assert not torch.equal(first, ro.actions)
    assert all(torch.isfinite(v).all() for v in ro.segments().values())
    assert abs(float(ro.actions.std()) - 1.0) < 0.05  # logstd = 0, means ~ 0.01 scale
    assert float(vec.actions.abs().max()) <= 1.0
    mean, logstd, value = p.forward_eval(ro.observations[5])
    lp = torch.distributions.Normal(mean, logstd.exp()).log_prob(ro.actions[5]).sum(1)
    assert torch.allclose(lp, ro.logprobs[5], atol=2e-3, rtol=1e-3)  # rollout default: TF32 GEMMs like the reference
    assert torch.allclose(value.flatten(), ro.values[5], atol=5e-3, rtol=5e-3)
    vec.close()
