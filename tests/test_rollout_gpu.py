"""On-device rollout (BASELINE.json configs[3]), two-kernel form: fused policy step + env step, K steps in one CUDA
graph (policy_impl="fused"; the one-kernel form is covered by tests/test_rollout_kernel_gpu.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_graph_rollout_equals_eager_rollout_deterministic_policy():
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import RaceVec
    n, K = 3000, 16
    torch.manual_seed(0)
    policy = DronePolicy().cuda()
    outs = []
    for graph in (False, True):
        vec = RaceVec(n, seed=5, math="strict", max_moves=30)
        vec.reset(5)
        ro = DeviceRollout(vec, policy, horizon=K, use_graph=graph, deterministic=True, policy_impl="fused")
        if not graph:
            for _ in range(2):  # the graph path runs two warm-up steps before capture
                ro._one_step(0)
        ro.collect()
        ro.collect()
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in ro.segments().items()} | {"state": vec.get_state(), "steps": vec.step_count})
        vec.close()
    a, b = outs
    assert a["steps"] == b["steps"] == 2 * K + 2
    for k in ("observations", "actions", "rewards", "terminals", "values", "logprobs"):
        assert torch.equal(a[k], b[k]), k
    assert np.array_equal(a["state"].view(np.uint32), b["state"].view(np.uint32))
    assert a["observations"].shape == (n, K, 29) and a["terminals"].sum() > 0


def test_stochastic_rollout_statistics_and_contract():
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import RaceVec
    n, K = 8192, 32
    torch.manual_seed(1)
    policy = DronePolicy().cuda()
    with torch.no_grad():
        policy.decoder_logstd.fill_(-0.5)
    vec = RaceVec(n, seed=2)
    vec.reset(2)
    ro = DeviceRollout(vec, policy, horizon=K, use_graph=True, policy_impl="fused")
    ro.collect()
    first = ro.actions.clone()
    ro.collect()
    torch.cuda.synchronize()
    assert not torch.equal(first, ro.actions)  # fresh noise on every replay
    seg = ro.segments()
    assert all(torch.isfinite(v).all() for v in seg.values())
    assert float(seg["rewards"].abs().max()) <= 1.0
    # actions ~ Normal(mean ~ 0, exp(-0.5)); log-prob consistent with the stored action
    std = float(ro.actions.std())
    assert abs(std - np.exp(-0.5)) < 0.02
    mean, logstd, _ = policy.forward_eval(ro.observations[3])
    lp = torch.distributions.Normal(mean, logstd.exp()).log_prob(ro.actions[3]).sum(1)
    assert torch.allclose(lp, ro.logprobs[3], atol=1e-4, rtol=1e-4)
    # the env saw the clipped action
    assert float(vec.actions.abs().max()) <= 1.0
    assert vec.step_count == 2 * K + 2
    vec.close()
