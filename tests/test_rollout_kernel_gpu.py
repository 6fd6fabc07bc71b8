"""The rollout as one kernel (b2d_race_rollout, csrc/rollout_kernels.cuh: tcgen05 policy GEMMs with TMEM
accumulators + the env step in registers, K steps per launch) against the oracle.

  * policy half: stored actions / values / log-probs of every step against oracle/policy_oracle.py evaluated on
    the observation the kernel stored for that step (TF32 operands, hidden activations truncated: what the
    kernel documents), tolerance 2e-4 like the two-kernel TF32 policy step;
  * env half: with math="strict" the transition (state, observation, reward, terminal after a step) given the
    stored action is bit-exact against the CPU restatement of the step, resets (Philox stream) included;
  * a K-step launch equals K one-step launches bit for bit (state carried in registers == state carried in HBM);
  * contract: buffers after the launch, step counter, episode statistics, noise advancing between launches.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_policy_cpu import _random_policy, _weights  # noqa: E402


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _rollout(n, K, seed=3, math="strict", policy=None, noise_seed=(7 << 32) | 5, deterministic=False, max_moves=1000):
    from drone_b200.rollout import DeviceRollout
    from drone_b200.vec import RaceVec
    vec = RaceVec(n, seed=seed, math=math, max_moves=max_moves)
    vec.reset(seed)
    p = policy if policy is not None else _random_policy().cuda()
    ro = DeviceRollout(vec, p, horizon=K, policy_impl="rollout_kernel", noise_seed=noise_seed, deterministic=deterministic)
    return vec, p, ro


@pytest.mark.parametrize("n", [4133, 128, 77])
def test_policy_outputs_match_the_oracle_every_step(n):
    from oracle import policy_oracle as pol
    K = 6
    vec, p, ro = _rollout(n, K, math="fast")
    rew0, term0 = vec.rewards.clone(), vec.terminals.clone()
    ro.collect()
    torch.cuda.synchronize()
    w = _weights(p)
    obs = ro.observations.cpu().numpy()
    for k in range(K):
        prev_rew = rew0.cpu().numpy() if k == 0 else None
        ref = pol.policy_act(w, obs[k], np.zeros(n, np.float32), np.zeros(n, np.uint8), call=k, seed=(7 << 32) | 5,
                             tf32_gemm=True, hidden="truncate")
        got_a, got_v, got_lp = ro.actions[k].cpu().numpy(), ro.values[k].cpu().numpy(), ro.logprobs[k].cpu().numpy()
        assert np.abs(got_v - ref["values"]).max() < 2e-4 * max(1.0, np.abs(ref["values"]).max()), k
        assert np.abs(got_a - ref["actions"]).max() < 2e-4 * max(1.0, np.abs(ref["actions"]).max()), k
        assert np.abs(got_lp - ref["logprobs"]).max() < 2e-4, k
        exact = pol.policy_act(w, obs[k], np.zeros(n, np.float32), np.zeros(n, np.uint8), call=k, seed=(7 << 32) | 5)
        assert np.abs(got_v - exact["values"]).max() < 5e-3 * max(1.0, np.abs(exact["values"]).max())
        assert np.abs(got_a - exact["actions"]).max() < 5e-3 * max(1.0, np.abs(exact["actions"]).max())
        if prev_rew is not None:
            assert np.array_equal(ro.rewards[0].cpu().numpy(), np.clip(prev_rew, -1, 1))
            assert np.array_equal(ro.terminals[0].cpu().numpy(), term0.cpu().numpy().astype(np.float32))
    assert int(ro.counter[0].item()) == K
    assert vec.step_count == K
    assert float(vec.actions.abs().max()) <= 1.0
    assert np.array_equal(vec.actions.cpu().numpy(), np.clip(ro.actions[K - 1].cpu().numpy(), -1, 1))
    vec.close()


def test_env_half_is_bit_exact_given_the_stored_actions(oracle):
    """strict env math: feed the CPU restatement the actions the kernel stored; every observation the kernel
    stored for the NEXT step, every reward / terminal and the final state must be identical."""
    n, K, seed = 3000, 24, 9
    vec, p, ro = _rollout(n, K, seed=seed, math="strict", max_moves=13)  # timeouts every 13 steps on top of OOB resets
    cpu = oracle.OrcRace(n, seed=seed, max_moves=13)
    cpu.reset(seed, mode=oracle.RESET_PHILOX)
    assert np.array_equal(_bits(vec.observations.cpu().numpy()), _bits(cpu.observations))
    ro.collect()
    ro_first = {k: v.clone() for k, v in ro.segments().items()}
    torch.cuda.synchronize()
    acts = ro.actions.cpu().numpy()
    obs = ro.observations.cpu().numpy()
    rew = ro.rewards.cpu().numpy()
    term = ro.terminals.cpu().numpy()
    nterm = 0
    for k in range(K):
        assert np.array_equal(_bits(obs[k]), _bits(cpu.observations)), f"observation of step {k}"
        cpu.step(np.clip(acts[k], -1, 1), mode=oracle.RESET_PHILOX)
        nterm += int(cpu.terminals.sum())
        if k + 1 < K:
            assert np.array_equal(rew[k + 1], np.clip(cpu.rewards, -1, 1)), f"reward after step {k}"
            assert np.array_equal(term[k + 1], cpu.terminals.astype(np.float32)), f"terminal after step {k}"
    assert nterm > n  # every env has finished at least once
    assert np.array_equal(_bits(vec.observations.cpu().numpy()), _bits(cpu.observations))
    assert np.array_equal(_bits(vec.rewards.cpu().numpy()), _bits(cpu.rewards))
    assert np.array_equal(vec.terminals.cpu().numpy(), cpu.terminals)
    assert np.array_equal(_bits(vec.get_state()), _bits(cpu.get_state()))
    got, ref = vec.log(), cpu.log()
    assert got["n"] == float(ref[8]) == float(nterm)
    assert got["episode_length"] == pytest.approx(ref[1] / ref[8], rel=1e-6)
    assert got["timeout"] == pytest.approx(ref[5] / ref[8], rel=1e-6)
    # a second collection continues where the first stopped (state, counters and noise all advanced on the device)
    ro.collect()
    torch.cuda.synchronize()
    assert vec.step_count == 2 * K
    assert not torch.equal(ro_first["actions"], ro.segments()["actions"])
    acts2 = ro.actions.cpu().numpy()
    for k in range(K):
        assert np.array_equal(_bits(ro.observations[k].cpu().numpy()), _bits(cpu.observations)), f"second collection, step {k}"
        cpu.step(np.clip(acts2[k], -1, 1), mode=oracle.RESET_PHILOX)
    assert np.array_equal(_bits(vec.get_state()), _bits(cpu.get_state()))
    vec.close()
    cpu.close()


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_one_launch_of_k_steps_equals_k_launches_of_one_step(math):
    n, K = 5000, 16
    p = _random_policy(seed=4).cuda()
    a_vec, _, a = _rollout(n, K, math=math, policy=p, max_moves=11)
    b_vec, _, b = _rollout(n, 1, math=math, policy=p, max_moves=11)
    a.collect()
    rows = {k: [] for k in ("observations", "actions", "logprobs", "rewards", "terminals", "values")}
    for _ in range(K):
        b.collect()
        for k in rows:
            rows[k].append(getattr(b, k)[0].clone())
    torch.cuda.synchronize()
    for k in rows:
        assert torch.equal(getattr(a, k).view(torch.int32), torch.stack(rows[k]).view(torch.int32)), k
    assert np.array_equal(_bits(a_vec.get_state()), _bits(b_vec.get_state()))
    assert torch.equal(a_vec.observations.view(torch.int32), b_vec.observations.view(torch.int32))
    assert a_vec.log() == b_vec.log()
    assert a_vec.step_count == b_vec.step_count == K
    a_vec.close()
    b_vec.close()


def test_fast_rollout_statistics_and_agreement_with_the_two_kernel_form():
    """math='fast' (what bench.py times): finite experience, Normal statistics of the samples, log-probs consistent
    with torch's own distribution on the stored observations, and the same first step as the two-kernel rollout."""
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import RaceVec
    n, K = 1 << 16, 32
    torch.manual_seed(1)
    policy = DronePolicy().cuda()
    with torch.no_grad():
        policy.decoder_logstd.fill_(-0.5)
    vec = RaceVec(n, seed=2)
    vec.reset(2)
    ro = DeviceRollout(vec, policy, horizon=K)
    assert ro.policy_impl == "rollout_kernel"
    ro.collect()
    first = ro.actions.clone()
    ro.collect()
    torch.cuda.synchronize()
    assert not torch.equal(first, ro.actions)
    seg = ro.segments()
    assert all(torch.isfinite(v).all() for v in seg.values())
    assert float(seg["rewards"].abs().max()) <= 1.0
    assert abs(float(ro.actions.std()) - np.exp(-0.5)) < 0.02
    mean, logstd, value = policy.forward_eval(ro.observations[5])
    lp = torch.distributions.Normal(mean, logstd.exp()).log_prob(ro.actions[5]).sum(1)
    assert torch.allclose(lp, ro.logprobs[5], atol=2e-3, rtol=1e-3)
    assert torch.allclose(value.flatten(), ro.values[5], atol=5e-3, rtol=5e-3)
    assert vec.step_count == 2 * K
    assert ro.kernel_launches >= 2
    stats = vec.log()
    assert stats["n"] > 0 and 0.0 <= stats["oob"] <= 1.0
    # the two-kernel form from the same start: identical first observation, first-step values within TF32 distance
    vec2 = RaceVec(n, seed=2)
    vec2.reset(2)
    ro2 = DeviceRollout(vec2, policy, horizon=4, policy_impl="fused", use_graph=False, deterministic=True)
    vec3 = RaceVec(n, seed=2)
    vec3.reset(2)
    ro3 = DeviceRollout(vec3, policy, horizon=4, policy_impl="rollout_kernel", deterministic=True)
    ro2.collect()
    ro3.collect()
    torch.cuda.synchronize()
    assert torch.equal(ro2.observations[0], ro3.observations[0])
    assert torch.allclose(ro2.values[0], ro3.values[0], atol=2e-3, rtol=2e-3)
    assert torch.allclose(ro2.actions[0], ro3.actions[0], atol=2e-3, rtol=2e-3)
    for v in (vec, vec2, vec3):
        v.close()


def test_argument_errors():
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import RaceVec, SwarmVec
    vec = RaceVec(256, seed=1)
    vec.reset(1)
    with pytest.raises(ValueError):
        DeviceRollout(vec, DronePolicy(hidden_size=64).cuda(), policy_impl="rollout_kernel")
    sw = SwarmVec(16, 8, 5, seed=1)
    with pytest.raises(ValueError):
        DeviceRollout(sw, DronePolicy(obs_dim=41).cuda(), policy_impl="rollout_kernel")
    sw.close()
    vec.close()
