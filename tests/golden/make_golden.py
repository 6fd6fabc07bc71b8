"""Generates the golden vectors in this directory from the UNMODIFIED reference C
(oracle/_ref, built by oracle/Makefile from /root/reference/pufferlib/pufferlib/ocean).

    python tests/golden/make_golden.py          # needs /root/reference (build container only)

The reference's own tests pin nothing for the drone envs, so these
files are the pin: every run records, from the reference itself,
  init_state / init_obs     full env state and observation rows after vec_reset(seed)
  tape                      the action tape (cycled t % 16; |a| up to 1.3 so clamping is exercised)
  term, rew                 every step's terminals and rewards
  obs_hash                  FNV-1a of each env's 29 observation words at every step (bit-exactness check)
  obs_steps / obs_full      complete observation rows at a few steps
  ev_t / ev_env / ev_blob   every auto-reset: when, which env, and the post-reset state blob
                            (= the payload a parity run injects in place of libc rand())
  clamped_hash              hash of the in-place clamped action buffer at every step
  final_state, log          state after the last step; float-wise summed Log (EB:572-580)
Blob layout: oracle/ref_shim_race.c / include/b200drone.h b2d_get_state.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from _util import action_tape, row_hash  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OBS_STEPS = [0, 1, 2, 3, 10, 50, 100, 250, 399]


def race_golden(n, T, seed, max_rings=10, max_moves=1000, tape_seed=1234, scale=1.3):
    env = po.RefRace(n, max_rings=max_rings, max_moves=max_moves)
    tape = action_tape(n, seed=tape_seed, scale=scale)
    env.reset(seed)
    g = dict(meta=np.array([n, T, seed, max_rings, max_moves], np.int64), tape=tape,
             init_state=env.get_state(), init_obs=env.observations.copy())
    term = np.zeros((T, n), np.uint8)
    rew = np.zeros((T, n), np.float32)
    oh = np.zeros((T, n), np.uint32)
    ch = np.zeros((T, n), np.uint32)
    ev_t, ev_env, ev_blob, obs_full, obs_steps = [], [], [], [], []
    for t in range(T):
        env.step(tape[t % len(tape)])
        term[t], rew[t] = env.terminals, env.rewards
        oh[t] = row_hash(env.observations)
        ch[t] = row_hash(env.actions)
        idx = np.flatnonzero(env.terminals)
        if len(idx):
            ev_t += [t] * len(idx)
            ev_env += idx.tolist()
            ev_blob.append(env.get_state(idx))
        if t in OBS_STEPS or t == T - 1:
            obs_steps.append(t)
            obs_full.append(env.observations.copy())
    g.update(term=term, rew=rew, obs_hash=oh, clamped_hash=ch, ev_t=np.array(ev_t, np.int32),
             ev_env=np.array(ev_env, np.int32),
             ev_blob=np.concatenate(ev_blob) if ev_blob else np.zeros((0, env.blob), np.float32),
             obs_steps=np.array(obs_steps, np.int32), obs_full=np.array(obs_full),
             final_state=env.get_state(), log=env.log())
    env.close()
    return g


def swarm_golden(n, A, T, seed, max_rings=5, tape_seed=77, scale=1.2):
    """Swarm vectors.  Outputs come from the unmodified reference (RefSwarm).  The results of its
    random draws (respawns, env-wide resets) are not observable from outside c_step, so they are
    taken from the CPU restatement run on the same libc stream right after, which must reproduce
    the reference's outputs bit for bit at every step before its draws are accepted."""
    tape = action_tape(n * A, seed=tape_seed, scale=scale)
    ref = po.RefSwarm(n, A, max_rings)
    ref.reset(seed)
    init_obs = ref.observations.copy()
    term = np.zeros((T, n * A), np.uint8)
    rew = np.zeros((T, n * A), np.float32)
    oh = np.zeros((T, n * A), np.uint32)
    obs_full, obs_steps = [], []
    for t in range(T):
        ref.step(tape[t % len(tape)])
        term[t], rew[t], oh[t] = ref.terminals, ref.rewards, row_hash(ref.observations)
        if t in OBS_STEPS or t in (1022, 1023, 1024) or t == T - 1:
            obs_steps.append(t)
            obs_full.append(ref.observations.copy())
    fin_env, fin_ag = ref.get_state()
    log = ref.log()
    ref.close()

    orc = po.OrcSwarm(n, A, max_rings)
    orc.reset(seed, mode=po.RESET_LIBC)
    assert np.array_equal(orc.observations.view(np.uint32), init_obs.view(np.uint32))
    pay0 = np.concatenate([orc.pay_agent.reshape(n, -1), orc.pay_env], axis=1)
    ev_t, ev_row, ev_pay, er_t, er_env, er_pay = [], [], [], [], [], []
    for t in range(T):
        orc.step(tape[t % len(tape)], mode=po.RESET_LIBC)
        assert np.array_equal(row_hash(orc.observations), oh[t]), f"restatement diverged from the reference at step {t}"
        assert np.array_equal(orc.rewards.view(np.uint32), rew[t].view(np.uint32)) and np.array_equal(orc.terminals, term[t])
        envs = np.flatnonzero(orc.flag_env)
        rows = np.flatnonzero(orc.flag_agent | np.repeat(orc.flag_env, A))
        if len(rows):
            ev_t += [t] * len(rows)
            ev_row += rows.tolist()
            ev_pay.append(orc.pay_agent[rows].copy())
        if len(envs):
            er_t += [t] * len(envs)
            er_env += envs.tolist()
            er_pay.append(orc.pay_env[envs].copy())
    e2, a2 = orc.get_state()
    assert np.array_equal(e2.view(np.uint32), fin_env.view(np.uint32)) and np.array_equal(a2.view(np.uint32), fin_ag.view(np.uint32))
    orc.close()
    return dict(meta=np.array([n, A, T, seed, max_rings], np.int64), tape=tape, init_obs=init_obs, init_payload=pay0,
                term=term, rew=rew, obs_hash=oh, obs_steps=np.array(obs_steps, np.int32), obs_full=np.array(obs_full),
                ev_t=np.array(ev_t, np.int32), ev_row=np.array(ev_row, np.int32),
                ev_pay=np.concatenate(ev_pay) if ev_pay else np.zeros((0, po.SWARM_AGENT_PAYLOAD), np.float32),
                er_t=np.array(er_t, np.int32), er_env=np.array(er_env, np.int32),
                er_pay=np.concatenate(er_pay) if er_pay else np.zeros((0, 2 + 6 * max_rings), np.float32),
                final_env=fin_env, final_agents=fin_ag, log=log)


if __name__ == "__main__":
    if not po.have_ref():
        po.build(quiet=False)
    assert po.have_ref(), "oracle/_ref is missing and /root/reference is not here to build it"
    for name, kw in [("race_n64_T1000_seed42.npz", dict(n=64, T=1000, seed=42)),
                     ("race_n48_T400_seed0.npz", dict(n=48, T=400, seed=0)),
                     # short episodes: every truncation path (max_moves) and tiny ring counts
                     ("race_n32_T300_seed7_moves25_rings3.npz", dict(n=32, T=300, seed=7, max_rings=3, max_moves=25, scale=0.6))]:
        g = race_golden(**kw)
        np.savez_compressed(os.path.join(HERE, name), **g)
        print(name, "resets:", len(g["ev_t"]), "bytes:", os.path.getsize(os.path.join(HERE, name)))
    for name, kw in [("swarm_n4_A8_T1100_seed3.npz", dict(n=4, A=8, T=1100, seed=3)),
                     ("swarm_n3_A16_T1060_seed12.npz", dict(n=3, A=16, T=1060, seed=12, max_rings=3)),
                     ("swarm_n6_A1_T1040_seed5.npz", dict(n=6, A=1, T=1040, seed=5)),
                     ("swarm_n2_A64_T300_seed21.npz", dict(n=2, A=64, T=300, seed=21))]:
        g = swarm_golden(**kw)
        np.savez_compressed(os.path.join(HERE, name), **g)
        print(name, "respawn rows:", len(g["ev_t"]), "env resets:", len(g["er_t"]), "tasks:", g["final_env"][:, 1],
              "bytes:", os.path.getsize(os.path.join(HERE, name)))
