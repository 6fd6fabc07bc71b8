"""CPU tests of the swarm oracle: the restatement (oracle/drone_oracle.c "swarm env") against the
golden vectors from the unmodified reference and against the reference itself (oracle/_ref).
Bar: bit-exact, every word, including the index-ordered neighbour visibility, per-agent
respawns and the env-wide reset every 1023 ticks."""
import glob
import os

import numpy as np
import pytest

from _util import GOLDEN_DIR, action_tape, bits, load_golden, row_hash

SWARM_GOLDEN = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "swarm_*.npz")))


def swarm_payload_at(g, t, n, A, R):
    """Payload rows of step t rebuilt from the golden file's sparse events."""
    pa = np.zeros((n * A, 41), np.float32)
    pe = np.zeros((n, 2 + 6 * R), np.float32)
    sel = g["ev_t"] == t
    pa[g["ev_row"][sel]] = g["ev_pay"][sel]
    sel = g["er_t"] == t
    pe[g["er_env"][sel]] = g["er_pay"][sel]
    return pa, pe


def test_swarm_golden_present():
    assert len(SWARM_GOLDEN) >= 4


@pytest.mark.parametrize("name", SWARM_GOLDEN)
def test_swarm_restatement_replays_reference_golden(oracle, name):
    g = load_golden(name)
    n, A, T, seed, R = (int(x) for x in g["meta"])
    env = oracle.OrcSwarm(n, A, R)
    W = A * 41
    env.pay_agent[:] = g["init_payload"][:, :W].reshape(n * A, 41)
    env.pay_env[:] = g["init_payload"][:, W:]
    env.reset(seed, mode=oracle.RESET_INJECT)
    assert np.array_equal(bits(env.observations), bits(g["init_obs"]))
    full = dict(zip(g["obs_steps"].tolist(), g["obs_full"]))
    for t in range(T):
        pa, pe = swarm_payload_at(g, t, n, A, R)
        env.pay_agent[:] = pa
        env.pay_env[:] = pe
        env.step(g["tape"][t % 16], mode=oracle.RESET_INJECT)
        assert np.array_equal(env.terminals, g["term"][t]), f"terminals differ at step {t}"
        assert np.array_equal(bits(env.rewards), bits(g["rew"][t])), f"rewards differ at step {t}"
        assert np.array_equal(row_hash(env.observations), g["obs_hash"][t]), f"observations differ at step {t}"
        if t in full:
            assert np.array_equal(bits(env.observations), bits(full[t]))
    e, a = env.get_state()
    assert np.array_equal(bits(e), bits(g["final_env"])) and np.array_equal(bits(a), bits(g["final_agents"]))
    assert np.array_equal(bits(env.log()), bits(g["log"]))
    env.close()


@pytest.mark.parametrize("n,A,R,T,seed", [(6, 8, 5, 1100, 3), (3, 1, 5, 1040, 1), (2, 64, 5, 120, 11), (5, 3, 2, 200, 9)])
def test_swarm_restatement_equals_reference_free_running(oracle, n, A, R, T, seed):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    tape = action_tape(n * A, scale=1.2)

    def run(env, **kw):
        env.reset(seed, **kw)
        h = [row_hash(env.observations)]
        for t in range(T):
            env.step(tape[t % 16], **kw)
            h.append(row_hash(env.observations, env.rewards, env.terminals))
        out = np.array(h), env.get_state(), env.log()
        env.close()
        return out

    a = run(oracle.RefSwarm(n, A, R))
    b = run(oracle.OrcSwarm(n, A, R), mode=oracle.RESET_LIBC)
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(bits(a[1][0]), bits(b[1][0])) and np.array_equal(bits(a[1][1]), bits(b[1][1]))
    assert np.array_equal(bits(a[2]), bits(b[2]))


def test_swarm_all_eight_tasks_match_reference(oracle):
    """Every task's target assignment (idle, hover, orbit, follow, cube, congo, flag, race) and
    the closed-form formation targets, against the reference."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    seen = set()
    for seed in range(6):
        n, A, R, T = 48, 5, 3, 12
        tape = action_tape(n * A, seed=seed, scale=0.8)
        ref = oracle.RefSwarm(n, A, R)
        ref.reset(seed)
        ro = [ref.observations.copy()]
        for t in range(T):
            ref.step(tape[t % 16])
            ro.append(ref.observations.copy())
        renv, rag = ref.get_state()
        ref.close()
        orc = oracle.OrcSwarm(n, A, R)
        orc.reset(seed, mode=oracle.RESET_LIBC)
        oo = [orc.observations.copy()]
        for t in range(T):
            orc.step(tape[t % 16], mode=oracle.RESET_LIBC)
            oo.append(orc.observations.copy())
        oenv, oag = orc.get_state()
        orc.close()
        assert np.array_equal(bits(np.array(ro)), bits(np.array(oo)))
        assert np.array_equal(bits(rag), bits(oag))
        seen |= set(int(x) for x in renv[:, 1])
    assert seen == set(range(8)), f"tasks covered: {sorted(seen)}"


def test_swarm_philox_stream_is_shard_invariant_and_sane(oracle):
    n, A, R, T = 8, 6, 4, 80
    tape = action_tape(n * A, scale=1.0)
    whole = oracle.OrcSwarm(n, A, R, seed=4)
    whole.reset(4, mode=oracle.RESET_PHILOX)
    hi = oracle.OrcSwarm(n // 2, A, R, seed=4, env_id_base=n // 2)
    hi.reset(4, mode=oracle.RESET_PHILOX)
    for t in range(T):
        whole.step(tape[t % 16], mode=oracle.RESET_PHILOX)
        hi.step(tape[t % 16][(n // 2) * A:], mode=oracle.RESET_PHILOX)
    assert np.array_equal(bits(whole.observations[(n // 2) * A:]), bits(hi.observations))
    env, ag = whole.get_state()
    assert np.all(np.abs(ag[..., 0]) <= 30.0 + 1e-3) and np.all(np.abs(ag[..., 2]) <= 10.0 + 1e-3)
    arm = ag[..., 21]
    assert arm.min() >= 0.05 - 1e-6 and arm.max() <= 0.2 + 1e-6  # size ~ U(0.1, 0.4)
    assert np.isfinite(whole.observations).all()
    whole.close()
    hi.close()
