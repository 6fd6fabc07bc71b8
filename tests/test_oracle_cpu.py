"""CPU tests of the oracle (test infrastructure): the restatement in oracle/drone_oracle.c
against (a) the golden vectors generated from the unmodified reference C
(tests/golden/make_golden.py) and (b) the reference itself when oracle/_ref is built.

The reference's own test-suite pins nothing for the drone envs, so
these two checks are what pins the oracle.  Bar: bit-exact, every word.
"""
import glob
import os

import numpy as np
import pytest

from _util import GOLDEN_DIR, action_tape, bits, load_golden, row_hash

RACE_GOLDEN = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "race_*.npz")))


def _payload_for_step(g, t, n, blob):
    pl = np.zeros((n, blob), np.float32)
    sel = g["ev_t"] == t
    pl[g["ev_env"][sel]] = g["ev_blob"][sel]
    return pl


def test_golden_files_present():
    assert len(RACE_GOLDEN) >= 3


@pytest.mark.parametrize("name", RACE_GOLDEN)
def test_restatement_reproduces_reference_golden(oracle, name):
    """Inject-mode replay: same initial state + action tape, the reference's post-reset
    states injected at every auto-reset.  Every output word must equal the reference's."""
    g = load_golden(name)
    n, T, seed, max_rings, max_moves = (int(x) for x in g["meta"])
    env = oracle.OrcRace(n, max_rings=max_rings, max_moves=max_moves)
    env.put_state(g["init_state"])
    env.observe()
    assert np.array_equal(bits(env.observations), bits(g["init_obs"]))
    full = dict(zip(g["obs_steps"].tolist(), g["obs_full"]))
    for t in range(T):
        env.step(g["tape"][t % 16], mode=oracle.RESET_INJECT, payload=_payload_for_step(g, t, n, env.blob))
        assert np.array_equal(env.terminals, g["term"][t]), f"terminals differ at step {t}"
        assert np.array_equal(bits(env.rewards), bits(g["rew"][t])), f"rewards differ at step {t}"
        assert np.array_equal(row_hash(env.observations), g["obs_hash"][t]), f"observations differ at step {t}"
        assert np.array_equal(row_hash(env.actions), g["clamped_hash"][t]), f"clamped actions differ at step {t}"
        if t in full:
            assert np.array_equal(bits(env.observations), bits(full[t]))
    assert np.array_equal(bits(env.get_state()), bits(g["final_state"]))
    assert np.array_equal(bits(env.log()), bits(g["log"]))
    env.close()


@pytest.mark.parametrize("n,T,seed,kw", [(256, 600, 42, {}), (97, 300, 3, dict(max_rings=2, max_moves=13))])
def test_restatement_equals_reference_free_running_libc(oracle, n, T, seed, kw):
    """No injection: both sides draw resets from libc rand() in the reference's draw order
    (runs are sequential because rand() is process-global).  Needs oracle/_ref."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    tape = action_tape(n)

    def run(env, **step_kw):
        env.reset(seed, **step_kw)
        h = [row_hash(env.observations)]
        for t in range(T):
            env.step(tape[t % 16], **step_kw)
            h.append(row_hash(env.observations, env.rewards, env.terminals, env.actions))
        out = np.array(h), env.get_state(), env.log()
        env.close()
        return out

    a = run(oracle.RefRace(n, **kw))
    b = run(oracle.OrcRace(n, **kw), mode=oracle.RESET_LIBC)
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(bits(a[1]), bits(b[1]))
    assert np.array_equal(bits(a[2]), bits(b[2]))
    assert a[2][8] > 0  # episodes did finish


def test_golden_matches_live_reference(oracle):
    """The committed golden file is what the reference build in this container produces."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    g = load_golden("race_n48_T400_seed0.npz")
    n, T, seed, max_rings, max_moves = (int(x) for x in g["meta"])
    env = oracle.RefRace(n, max_rings=max_rings, max_moves=max_moves)
    env.reset(seed)
    assert np.array_equal(bits(env.get_state()), bits(g["init_state"]))
    for t in range(T):
        env.step(g["tape"][t % 16])
        assert np.array_equal(row_hash(env.observations), g["obs_hash"][t])
    env.close()


def test_reference_quirks_pinned_by_golden():
    """Behaviours a 'cleaned up' implementation would lose."""
    g = load_golden("race_n64_T1000_seed42.npz")
    obs = g["obs_full"]
    assert np.array_equal(bits(obs[..., 6:12]), bits(obs[..., 0:6]))  # next ring == current ring (drone_race.h:77)
    assert set(np.unique(g["rew"]).tolist()) <= {-1.0, 0.0, 1.0}
    # log.score is zeroed at the top of every step (:160): only episodes that ended in the very
    # last step can contribute, so the summed score is tiny next to the episode count
    assert g["log"][6] <= g["term"][-1].sum() * 10
    # observation rows of terminated envs are the POST-reset observation (zero velocity)
    k = list(g["obs_steps"]).index(100)
    done = g["term"][100] == 1
    if done.any():
        assert np.all(obs[k][done][:, 12:18] == 0.0)


# ---- the device reset stream, restated on the CPU ------------------------------------------
def test_philox_known_answers(oracle):
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_deterministic_sincos_accuracy(oracle):
    th = np.linspace(0.0, 2.0 * np.pi, 4001).astype(np.float32)
    got = np.array([oracle.sincos_det(float(t)) for t in th])
    assert np.abs(got[:, 0] - np.sin(th.astype(np.float64))).max() < 2e-7
    assert np.abs(got[:, 1] - np.cos(th.astype(np.float64))).max() < 2e-7


def test_philox_reset_distribution_matches_reference_reset(oracle):
    """Device-native resets are validated distributionally against the reference's c_reset:
    same supports and constraints, same means within sampling error."""
    n = 20000
    orc = oracle.OrcRace(n, seed=5)
    orc.reset(5, mode=oracle.RESET_PHILOX)
    a = orc.get_state()
    rings = a[:, 33:].reshape(n, 10, 6)
    # constraints of c_reset (drone_race.h:127-151, dronelib.h:451-460)
    assert np.all(np.abs(rings[..., :3]) <= 6.0)
    d = np.linalg.norm(np.diff(rings[..., :3], axis=1), axis=2)
    assert (d >= 4.0 - 1e-5).mean() > 0.9999  # 16 bounded attempts instead of an unbounded loop
    assert np.allclose(np.linalg.norm(rings[..., 3:], axis=2), 1.0, atol=1e-5)
    spawn = a[:, 0:3]
    assert np.all(np.abs(spawn) <= 9.0)
    assert (np.linalg.norm(spawn - rings[:, 0, :3], axis=1) >= 4.0 - 1e-5).mean() > 0.9999
    assert np.all(a[:, 3:6] == 0) and np.all(a[:, 6] == 1) and np.all(a[:, 7:17] == 0)
    arm = a[:, 21]
    assert arm.min() >= 0.025 - 1e-6 and arm.max() <= 0.4 + 1e-6
    assert np.all(np.abs(a[:, 26] / 9.81 - 1.0) <= 0.01 + 1e-6)
    assert np.all(np.abs(a[:, 28] / 0.1 - 1.0) <= 0.1 + 1e-5)
    if oracle.have_ref():
        ref = oracle.RefRace(n)
        ref.reset(5)
        b = ref.get_state()
        ref.close()
        for col in (17, 21, 22, 26, 27, 28, 29):  # mass, arm, k_thrust, gravity, max_rpm, k_mot, j_mot
            sa, sb = a[:, col].astype(np.float64), b[:, col].astype(np.float64)
            se = sb.std() / np.sqrt(n) * 6 + 1e-12
            assert abs(sa.mean() - sb.mean()) < se * 2, f"param column {col}: {sa.mean()} vs {sb.mean()}"
        assert abs(np.abs(a[:, 0:3]).mean() - np.abs(b[:, 0:3]).mean()) < 0.1
    orc.close()


def test_philox_episodes_are_pure_functions_of_seed_env_episode(oracle):
    """Shard invariance: env g of a 2-way split == env g of the unsplit run (env_id_base)."""
    n, T = 64, 200
    tape = action_tape(n, scale=1.0)
    whole = oracle.OrcRace(n, seed=9)
    whole.reset(9, mode=oracle.RESET_PHILOX)
    lo = oracle.OrcRace(n // 2, seed=9, env_id_base=0)
    hi = oracle.OrcRace(n // 2, seed=9, env_id_base=n // 2)
    lo.reset(9, mode=oracle.RESET_PHILOX)
    hi.reset(9, mode=oracle.RESET_PHILOX)
    for t in range(T):
        whole.step(tape[t % 16], mode=oracle.RESET_PHILOX)
        lo.step(tape[t % 16][: n // 2], mode=oracle.RESET_PHILOX)
        hi.step(tape[t % 16][n // 2:], mode=oracle.RESET_PHILOX)
    assert np.array_equal(bits(whole.observations), bits(np.concatenate([lo.observations, hi.observations])))
    assert np.allclose(whole.log(), lo.log() + hi.log(), rtol=1e-6)
    for e in (whole, lo, hi):
        e.close()


def test_edge_cases_single_env_and_zero_actions(oracle):
    env = oracle.OrcRace(1, max_rings=1, max_moves=3)
    env.reset(0, mode=oracle.RESET_PHILOX)
    terms = []
    for _ in range(7):
        env.step(np.zeros((1, 4), np.float32), mode=oracle.RESET_PHILOX)
        terms.append(int(env.terminals[0]))
    assert terms[2] == 1 and terms[5] == 1  # truncation every max_moves steps
    assert np.isfinite(env.observations).all()
    env.close()
