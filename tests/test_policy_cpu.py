"""CPU checks of the policy-step oracle (oracle/policy_oracle.py) against what the reference calls:
torch's own nn.Linear / nn.GELU / torch.distributions.Normal (pufferlib/models.py:41-98,
pufferlib/pytorch.py:189-199), and of its Philox restatement against the C oracle / Random123
known-answer vectors.  No compute call into the CUDA library is made here."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _weights(policy):
    g = lambda t: t.detach().cpu().numpy()  # noqa: E731
    return dict(encoder_weight=g(policy.encoder[0].weight), encoder_bias=g(policy.encoder[0].bias),
                decoder_mean_weight=g(policy.decoder_mean.weight), decoder_mean_bias=g(policy.decoder_mean.bias),
                decoder_logstd=g(policy.decoder_logstd), value_weight=g(policy.value.weight), value_bias=g(policy.value.bias))


def _random_policy(obs_dim=29, hidden=128, seed=0):
    from drone_b200.rollout import DronePolicy
    torch.manual_seed(seed)
    p = DronePolicy(obs_dim=obs_dim, hidden_size=hidden)
    with torch.no_grad():  # non-trivial biases, heads and log-std so every term is exercised
        for t in p.parameters():
            if t.dim() == 1 or t is p.decoder_logstd:
                t.uniform_(-0.5, 0.5)
        p.decoder_mean.weight.mul_(30.0)
    return p


def _forward64(p, obs):
    """Default.forward_eval (models.py:65-98) on a .double() module: same torch modules, float64 arithmetic."""
    hidden = p.encoder(obs.double())
    mean = p.decoder_mean(hidden)
    return mean, p.decoder_logstd.expand_as(mean), p.value(hidden)


def test_vectorised_philox_matches_the_c_oracle_and_random123_vectors(oracle):
    from oracle import policy_oracle as pol
    # Random123 known-answer vectors (kat_vectors: philox4x32 10)
    out = pol.philox4x32_10(np.uint32(0xFFFFFFFF), np.uint32(0xFFFFFFFF), np.uint32(0xFFFFFFFF), np.uint32(0xFFFFFFFF),
                            0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = pol.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(x) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    key = (0x9E3779B9, 0x12345678)
    got = np.stack(pol.philox4x32_10(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], *key), axis=1)
    for i in range(64):
        assert list(got[i]) == list(oracle.philox4x32_10(ctr[i], key))


def test_noise_is_standard_normal_and_keyed_by_row_call_seed():
    from oracle import policy_oracle as pol
    z = pol.noise(200000, call=3, seed=11)
    assert z.shape == (200000, 4)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs((z ** 4).mean() - 3.0) < 0.1  # kurtosis of a normal
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.01
    assert np.array_equal(z[100:200], pol.noise(100, 3, 11, row_id_base=100))  # shard-invariant
    assert not np.array_equal(z[:100], pol.noise(100, 4, 11))
    assert not np.array_equal(z[:100], pol.noise(100, 3, 12))


@pytest.mark.parametrize("obs_dim,hidden", [(29, 128), (41, 64)])
def test_oracle_forward_equals_torch_modules(obs_dim, hidden):
    from oracle import policy_oracle as pol
    p = _random_policy(obs_dim, hidden).double()
    obs = torch.randn(257, obs_dim) * 1.5
    mean, logstd, value = _forward64(p, obs)
    m, ls, v = pol.forward_eval(_weights(p), obs.numpy())
    assert np.allclose(m, mean.detach().numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(v, value.detach().numpy()[:, 0], rtol=1e-12, atol=1e-12)
    assert np.allclose(ls, logstd.detach().numpy())


def test_oracle_sampling_and_logprob_equal_torch_normal():
    from oracle import policy_oracle as pol
    p = _random_policy().double()
    w = _weights(p)
    rng = np.random.default_rng(1)
    obs = rng.normal(size=(500, 29)).astype(np.float32)
    rew = rng.normal(size=500).astype(np.float32) * 2
    term = (rng.random(500) < 0.1).astype(np.uint8)
    out = pol.policy_act(w, obs, rew, term, call=5, seed=9)
    mean, logstd, _ = _forward64(p, torch.from_numpy(obs))
    dist = torch.distributions.Normal(mean, torch.exp(logstd))
    lp = dist.log_prob(torch.from_numpy(out["actions"])).sum(1)
    assert np.allclose(out["logprobs"], lp.detach().numpy(), rtol=1e-10, atol=1e-10)
    assert np.allclose(out["actions"], (mean + torch.exp(logstd) * torch.from_numpy(out["noise"])).detach().numpy())
    assert out["rewards"].min() >= -1 and out["rewards"].max() <= 1 and np.array_equal(out["terminals"], term)
    assert np.abs(out["env_actions"]).max() <= 1.0
    det = pol.policy_act(w, obs, rew, term, call=5, seed=9, deterministic=True)
    assert np.allclose(det["actions"], mean.detach().numpy())


def test_tf32_rounding_is_cvt_rna():
    """round to nearest on the 10-bit mantissa, ties away from zero (PTX cvt.rna.tf32.f32)"""
    from oracle import policy_oracle as pol
    x = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 - 2.0 ** -23, -(1.0 + 2.0 ** -11), 1.0 + 2.0 ** -10, 3.14159265, 0.0], np.float32)
    want = np.array([1.0, 1.0 + 2.0 ** -10, 1.0, -(1.0 + 2.0 ** -10), 1.0 + 2.0 ** -10, 3.140625, 0.0], np.float32)
    assert np.array_equal(pol.tf32(x), want)
    r = np.random.default_rng(0).normal(size=100000).astype(np.float32)
    t = pol.tf32(r)
    assert (t.view(np.uint32) & 0x1FFF).max() == 0 and np.abs(t - r).max() <= np.abs(r).max() * 2.0 ** -11
    p = _random_policy()
    obs = r[:29 * 300].reshape(300, 29)
    m0, _, v0 = pol.forward_eval(_weights(p), obs)
    m1, _, v1 = pol.forward_eval(_weights(p), obs, tf32_gemm=True)
    assert 1e-6 < np.abs(v1 - v0).max() < 5e-3 and np.abs(m1 - m0).max() < 5e-3
