"""Edge cases on the GPU against the oracle (strict math, bit-exact): tiny and ragged vectors,
non-finite and out-of-range actions (the reference's clampf lets NaN through, dronelib.h:73-79),
smallest legal configurations, vec_log with nothing finished, state hooks on subsets."""
import numpy as np
import pytest

from _util import action_tape, bits

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 129, 1025])
def test_race_tiny_and_ragged_vectors(oracle, n):
    from drone_b200.vec import RaceVec
    seed = 3
    cpu = oracle.OrcRace(n, max_moves=17, seed=seed)
    cpu.reset(seed, mode=oracle.RESET_PHILOX)
    vec = RaceVec(n, max_moves=17, math="strict", seed=seed)
    vec.reset(seed)
    tape = action_tape(n, scale=1.0)
    dtape = torch.from_numpy(tape).cuda()
    for t in range(60):
        cpu.step(tape[t % 16], mode=oracle.RESET_PHILOX)
        vec.step(dtape[t % 16])
    torch.cuda.synchronize()
    assert np.array_equal(bits(vec.observations.cpu().numpy()), bits(cpu.observations))
    assert np.array_equal(vec.terminals.cpu().numpy(), cpu.terminals)
    assert np.array_equal(bits(vec.get_state()), bits(cpu.get_state()))
    vec.close()
    cpu.close()


def test_race_non_finite_and_huge_actions_propagate_like_the_reference(oracle):
    """NaN passes both clampf tests, +-inf and 1e30 clamp to +-1: whatever the reference's
    arithmetic does with them afterwards (NaN state, no OOB: comparisons are false) must come out
    of the kernel bit for bit, and must not disturb the other envs."""
    from drone_b200 import capi
    from drone_b200.vec import RaceVec
    n, seed = 256, 9
    cpu, kind = (oracle.RefRace(n), "reference") if oracle.have_ref() else (oracle.OrcRace(n), "port")
    cpu.reset(seed)
    vec = RaceVec(n, math="strict", write_clamped_actions=True)
    vec.set_reset_mode(capi.RESET_INJECT)
    vec.put_state(cpu.get_state())
    tape = action_tape(n, scale=1.0)
    tape[:, 5, 0] = np.nan
    tape[:, 9, :] = np.inf
    tape[:, 11, 2] = -np.inf
    tape[:, 17, 1] = 1e30
    tape[3, 40, 3] = np.nan  # turns NaN mid-episode
    dtape = torch.from_numpy(tape).cuda()
    for t in range(48):
        cpu.step(tape[t % 16])
        idx = np.flatnonzero(cpu.terminals)
        payload = np.zeros((n, cpu.blob), np.float32)
        if len(idx):
            payload[idx] = cpu.get_state(idx)
        vec.set_reset_payload(payload)
        vec.actions.copy_(dtape[t % 16])
        vec.step()
        obs = vec.observations.cpu().numpy()
        # NaN payloads differ in sign/payload bits between x86 and the GPU: compare NaN-ness, then bits elsewhere
        nan_ref, nan_dev = np.isnan(cpu.observations), np.isnan(obs)
        assert np.array_equal(nan_ref, nan_dev), f"NaN pattern differs at step {t} ({kind})"
        assert np.array_equal(bits(obs)[~nan_ref], bits(cpu.observations)[~nan_ref]), f"obs differ at step {t}"
        assert np.array_equal(vec.terminals.cpu().numpy(), cpu.terminals)
        act = vec.actions.cpu().numpy()
        assert np.array_equal(np.isnan(act), np.isnan(cpu.actions))
        assert np.array_equal(bits(act)[~np.isnan(act)], bits(cpu.actions)[~np.isnan(act)])
    assert np.isnan(cpu.observations[5]).any() and not np.isnan(cpu.observations[6]).any()
    vec.close()
    cpu.close()


def test_race_smallest_configuration_and_empty_log(oracle):
    from drone_b200.vec import RaceVec
    vec = RaceVec(1, max_rings=1, max_moves=1000, math="strict", seed=1)
    vec.reset(1)
    assert vec.log() == {}  # nothing finished: the reference returns {} (env_binding.h:582-585)
    cpu = oracle.OrcRace(1, max_rings=1, max_moves=1000, seed=1)
    cpu.reset(1, mode=oracle.RESET_PHILOX)
    z = np.zeros((1, 4), np.float32)
    for _ in range(5):
        cpu.step(z, mode=oracle.RESET_PHILOX)
        vec.step(torch.zeros((1, 4), device="cuda"))
    assert np.array_equal(bits(vec.observations.cpu().numpy()), bits(cpu.observations))
    vec.close()
    cpu.close()


def test_state_hooks_on_env_subsets_and_errors():
    from drone_b200.vec import RaceVec
    n = 300
    vec = RaceVec(n, math="strict", seed=4)
    vec.reset(4)
    full = vec.get_state()
    ids = [7, 0, 299, 123]
    sub = vec.get_state(ids)
    assert np.array_equal(bits(sub), bits(full[ids]))
    edit = sub.copy()
    edit[:, 0:3] = [[1.0, 2.0, 3.0]] * 4
    vec.put_state(edit, ids)
    back = vec.get_state()
    assert np.array_equal(bits(back[ids][:, :33]), bits(edit[:, :33]))
    untouched = np.setdiff1d(np.arange(n), ids)
    assert np.array_equal(bits(back[untouched]), bits(full[untouched]))
    with pytest.raises(ValueError):
        vec.get_state([n])
    with pytest.raises(ValueError):
        vec.step(torch.zeros((n, 4), device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError):
        vec.step(torch.zeros((n - 1, 4), device="cuda"))
    vec.close()


@pytest.mark.parametrize("E,A", [(1, 1), (1, 2), (3, 128), (7, 5), (130, 1)])
def test_swarm_tiny_and_ragged_vectors(oracle, E, A):
    from drone_b200.vec import SwarmVec
    seed, R = 6, 2
    cpu = oracle.OrcSwarm(E, A, R, seed=seed)
    cpu.reset(seed, mode=oracle.RESET_PHILOX)
    vec = SwarmVec(E, A, R, math="strict", seed=seed)
    vec.reset(seed)
    tape = action_tape(E * A, scale=1.1)
    dtape = torch.from_numpy(tape).cuda()
    for t in range(50):
        cpu.step(tape[t % 16], mode=oracle.RESET_PHILOX)
        vec.step(dtape[t % 16])
    torch.cuda.synchronize()
    assert np.array_equal(bits(vec.observations.cpu().numpy()), bits(cpu.observations))
    assert np.array_equal(bits(vec.rewards.cpu().numpy()), bits(cpu.rewards))
    env, ag = vec.split_state(vec.get_state())
    oenv, oag = cpu.get_state()
    assert np.array_equal(bits(env), bits(oenv)) and np.array_equal(bits(ag), bits(oag))
    vec.close()
    cpu.close()
