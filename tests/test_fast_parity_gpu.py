"""Parity of the PRODUCTION path -- math="fast", the kernels bench.py times -- at BASELINE.json's shapes.

north_star: "integer and boolean outputs (terminals, gate indices, reset events) must match
bit-exactly, continuous state / observations / rewards within a stated FP32 tolerance (1e-5
relative per step, drift bound reported over 1000 steps)".

Protocol.  The full-size vector runs on the device with its own counter-based reset
stream; slices of it -- start, an unaligned middle, the ragged tail -- are mirrored by the CPU
oracle created with env_id_base = slice start (same Philox stream, so resets need no injection).
Before EVERY step the slices' device state is overwritten with the oracle's exact state (put_state:
"per-step resync"), both sides step on the same actions, and then

  * integer / event outputs must be IDENTICAL: terminals, rewards that are event values (race:
    -1 / 0 / +1), tick, ring index, episodic return (race); terminals, ring index, episode length,
    collision count (swarm) -- the near-threshold guard of the fast kernels (race_strict_replay,
    swarm guard) exists to make this hold;
  * continuous outputs must lie within REL_TOL = 1e-5 relative, ABS_TOL = 1e-6 absolute.

The measured numbers (env-steps compared, flips, worst errors, guard replays, 1000-step free-running
drift) are written to gpurun_out/parity_r02.json; the copy committed as profiles/parity_r02.json is
what bench.py quotes in its "parity" key.
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("B2D_PARITY_OUT", os.path.join(ROOT, "gpurun_out", "parity_r02.json"))
REL_TOL, ABS_TOL = 1e-5, 1e-6
RACE_N = 1 << 20                      # BASELINE.json configs[1]
SWARM_ENVS, SWARM_A = 1 << 16, 64     # BASELINE.json configs[2]
_STATE_GROUPS = [(0, 3), (3, 6), (6, 10), (10, 13), (13, 17)]  # pos, vel, quat, omega, rpm


def _record(key, value):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    data = {}
    if os.path.exists(OUT):
        try:
            data = json.load(open(OUT))
        except Exception:  # noqa: BLE001
            data = {}
    data[key] = value
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _tol_err(got, ref, scale=None):
    """|got - ref| in units of the tolerance ABS_TOL + REL_TOL * scale (<= 1 means inside)."""
    ref = ref.astype(np.float64)
    scale = np.abs(ref) if scale is None else scale
    return np.abs(got.astype(np.float64) - ref) / (ABS_TOL + REL_TOL * scale)


def _state_err(got, ref):
    """State vectors: every component against the max-norm of the physical vector it belongs to."""
    worst = 0.0
    for lo, hi in _STATE_GROUPS:
        scale = np.abs(ref[:, lo:hi]).max(axis=1, keepdims=True)
        if got.shape[0]:
            worst = max(worst, float(_tol_err(got[:, lo:hi], ref[:, lo:hi], scale).max()))
    return worst


def _bench_tape(n, steps=16, seed=1234):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.rand((steps, n, 4), generator=g) * 2.0 - 1.0  # bench.py's action tape recipe


def test_race_fast_full_size_per_step_resync(oracle):
    from drone_b200.vec import RaceVec
    n, seed, T = RACE_N, 0, 110
    slices = [(0, 4096), (517_123, 517_123 + 4099), (n - 4097, n)]
    tape = _bench_tape(n)
    dtape, htape = tape.cuda(), tape.numpy()
    vec = RaceVec(n, max_rings=10, max_moves=1000, math="fast", seed=seed)
    vec.reset(seed)
    cpus = []
    for a, b in slices:
        cpu = oracle.OrcRace(b - a, max_rings=10, max_moves=1000, seed=seed, env_id_base=a)
        cpu.reset(seed, mode=oracle.RESET_PHILOX)
        cpus.append(cpu)
    ids = np.concatenate([np.arange(a, b) for a, b in slices])
    # let the episodes spread out first (both sides strict-equal at reset; 40 un-synced warm-up steps
    # on the oracle side only would desync the Philox episode numbers, so the warm-up is synced too)
    stats = dict(env_steps=0, terminal_flips=0, reward_flips=0, tick_flips=0, ring_idx_flips=0, return_flips=0,
                 terminals=0, ring_passes=0, collisions=0, worst_obs_tol=0.0, worst_state_tol=0.0,
                 worst_obs_abs=0.0, worst_pos_abs=0.0)
    for t in range(T):
        before = np.concatenate([cpu.get_state() for cpu in cpus])
        vec.put_state(before, env_ids=ids)
        for (a, b), cpu in zip(slices, cpus):
            cpu.step(htape[t % 16, a:b], mode=oracle.RESET_PHILOX)
        vec.step(dtape[t % 16])
        term = vec.terminals.cpu().numpy()[ids]
        rew = vec.rewards.cpu().numpy()[ids]
        obs = vec.observations.cpu().numpy()[ids]
        st = vec.get_state(ids)
        rterm = np.concatenate([c.terminals for c in cpus])
        rrew = np.concatenate([c.rewards for c in cpus])
        robs = np.concatenate([c.observations for c in cpus])
        rst = np.concatenate([c.get_state() for c in cpus])
        stats["env_steps"] += len(ids)
        stats["terminal_flips"] += int((term != rterm).sum())
        stats["reward_flips"] += int((rew.view(np.uint32) != rrew.view(np.uint32)).sum())
        stats["tick_flips"] += int((st[:, 30] != rst[:, 30]).sum())
        stats["ring_idx_flips"] += int((st[:, 31] != rst[:, 31]).sum())
        stats["return_flips"] += int((st[:, 32] != rst[:, 32]).sum())
        stats["terminals"] += int(rterm.sum())
        stats["ring_passes"] += int((rrew > 0).sum())
        stats["collisions"] += int(((rrew < 0) & (rterm == 1) & (np.abs(before[:, 0:3]).max(axis=1) <= 10)).sum())
        same = term == rterm
        stats["worst_obs_tol"] = max(stats["worst_obs_tol"], float(_tol_err(obs[same], robs[same]).max()))
        stats["worst_obs_abs"] = max(stats["worst_obs_abs"], float(np.abs(obs[same] - robs[same]).max()))
        live = same & (rterm == 0)  # finished envs hold the (identical) first state of their next episode
        stats["worst_state_tol"] = max(stats["worst_state_tol"], _state_err(st[live, :17], rst[live, :17]))
        stats["worst_pos_abs"] = max(stats["worst_pos_abs"], float(np.abs(st[live, 0:3] - rst[live, 0:3]).max()))
    stats["guard_replays_whole_vector"] = vec.guard_replays
    stats["guard_replay_rate"] = stats["guard_replays_whole_vector"] / float(n * T)
    stats["config"] = f"RaceVec({n}, math='fast'), slices {slices}, {T} steps, bench.py action tape, oracle = CPU restatement (Philox resets)"
    stats["tolerance"] = {"rel": REL_TOL, "abs": ABS_TOL, "note": "worst_*_tol are in units of abs + rel*|ref| (<= 1 passes)"}
    _record("race_fast_per_step", stats)
    for cpu in cpus:
        cpu.close()
    vec.close()
    assert stats["env_steps"] >= 1_000_000
    assert stats["terminals"] > 10_000 and stats["ring_passes"] > 0
    for k in ("terminal_flips", "reward_flips", "tick_flips", "ring_idx_flips", "return_flips"):
        assert stats[k] == 0, f"{k} = {stats[k]} over {stats['env_steps']} env-steps (integer outputs must be identical)"
    assert stats["worst_obs_tol"] <= 1.0, stats
    assert stats["worst_state_tol"] <= 1.0, stats


@pytest.mark.parametrize("recipe", ["bench_uniform", "near_hover"])
def test_race_fast_free_running_drift_1000_steps(oracle, recipe):
    """Drift without resync: device (fast) and oracle free-run from one state on the same actions with the
    same (Philox) resets; an env is compared for as long as its event history agrees with the oracle's."""
    from drone_b200.vec import RaceVec
    n, T, seed = 8192, 1000, 5
    if recipe == "bench_uniform":
        tape = _bench_tape(n).numpy()
    else:  # long episodes: the worst case for accumulated drift
        rng = np.random.default_rng(3)
        tape = (-0.24 + 0.05 * rng.standard_normal((16, n, 4))).astype(np.float32)
    cpu = oracle.OrcRace(n, seed=seed)
    cpu.reset(seed, mode=oracle.RESET_PHILOX)
    vec = RaceVec(n, math="fast", seed=seed)
    vec.reset(seed)
    dtape = torch.from_numpy(tape).cuda()
    agree = np.ones(n, bool)
    max_pos = max_quat = 0.0
    mean_pos, mean_quat, samples = [], [], []
    age = np.zeros(n, np.int64)  # steps since the env's last reset
    max_age_seen = 0
    for t in range(T):
        cpu.step(tape[t % 16], mode=oracle.RESET_PHILOX)
        vec.step(dtape[t % 16])
        term = vec.terminals.cpu().numpy()
        agree &= term == cpu.terminals
        age = np.where(cpu.terminals == 1, 0, age + 1)
        if t % 25 == 24 or t == T - 1:
            st, rst = vec.get_state(), cpu.get_state()
            a = agree
            dp = np.abs(st[a, 0:3] - rst[a, 0:3]).max(axis=1)
            dq = np.abs(st[a, 6:10] - rst[a, 6:10]).max(axis=1)
            max_pos, max_quat = max(max_pos, float(dp.max())), max(max_quat, float(dq.max()))
            mean_pos.append(float(dp.mean()))
            mean_quat.append(float(dq.mean()))
            max_age_seen = max(max_age_seen, int(age[a].max()))
            samples.append(int(a.sum()))
    res = dict(envs=n, steps=T, actions=recipe, max_abs_dpos_m=max_pos, mean_abs_dpos_m=float(np.mean(mean_pos)),
               max_abs_dquat=max_quat, mean_abs_dquat=float(np.mean(mean_quat)),
               envs_with_identical_event_history_after_1000_steps=float(agree.mean()),
               longest_episode_age_compared_steps=max_age_seen,
               note="no resync; compared every 25 steps over the envs whose terminal history still equals the oracle's")
    _record(f"race_fast_drift_{recipe}", res)
    vec.close()
    cpu.close()
    assert agree.mean() > 0.95, res
    assert max_pos < 5e-2 and max_quat < 5e-2, res


@pytest.mark.parametrize("A", [64, 16])
def test_swarm_fast_full_size_per_step_resync(oracle, A):
    from drone_b200.vec import SwarmVec
    n, R, seed, T = SWARM_ENVS, 10, 3, 104
    per = 5504 // A  # 5,504 drones per slice, 16,512 per step
    slices = [(0, per), (30_011, 30_011 + per), (n - per, n)]
    rows = n * A
    g = torch.Generator(device="cpu").manual_seed(99)
    tape = torch.rand((4, rows, 4), generator=g) * 2.0 - 1.0
    dtape, htape = tape.cuda(), tape.numpy()
    vec = SwarmVec(n, A, R, math="fast", seed=seed)
    vec.reset(seed)
    orcs = []
    for a, b in slices:
        o = oracle.OrcSwarm(b - a, A, R, seed=seed, env_id_base=a)
        o.reset(seed, mode=oracle.RESET_PHILOX)
        orcs.append(o)
    ids = np.concatenate([np.arange(a, b) for a, b in slices])
    rid = (ids[:, None] * A + np.arange(A)[None, :]).reshape(-1)
    stats = dict(drone_steps=0, terminal_flips=0, ring_idx_flips=0, episode_length_flips=0, collision_count_flips=0,
                 nearest_neighbour_flips=0, terminals=0, collisions=0, worst_obs_tol=0.0, worst_reward_abs=0.0,
                 worst_state_tol=0.0)
    grid = np.array([30.0, 30.0, 10.0])
    for t in range(T):
        before = [o.get_state() for o in orcs]
        vec.put_state(np.concatenate([vec.join_state(env, ag) for env, ag in before]), env_ids=ids)
        ref_col0 = np.concatenate([ag[:, :, 43].reshape(-1) for _, ag in before])
        for (a, b), o in zip(slices, orcs):
            o.step(htape[t % 4, a * A:b * A], mode=oracle.RESET_PHILOX)
        vec.step(dtape[t % 4])
        term = vec.terminals.cpu().numpy()[rid]
        rew = vec.rewards.cpu().numpy()[rid]
        obs = vec.observations.cpu().numpy()[rid]
        _, ag = vec.split_state(vec.get_state(ids))
        ag = ag.reshape(-1, 47)
        rterm = np.concatenate([o.terminals for o in orcs])
        rrew = np.concatenate([o.rewards for o in orcs])
        robs = np.concatenate([o.observations for o in orcs])
        rag = np.concatenate([o.get_state()[1].reshape(-1, 47) for o in orcs])
        stats["drone_steps"] += len(rid)
        stats["terminal_flips"] += int((term != rterm).sum())
        stats["ring_idx_flips"] += int((ag[:, 46] != rag[:, 46]).sum())
        stats["episode_length_flips"] += int((ag[:, 44] != rag[:, 44]).sum())
        stats["collision_count_flips"] += int((ag[:, 43] != rag[:, 43]).sum())
        stats["terminals"] += int(rterm.sum())
        stats["collisions"] += int((rag[:, 43] > ref_col0).sum())
        # Observation columns that are DIFFERENCES of positions (clamp(target - pos), (target - pos) / GRID,
        # clamp(nearest - pos), body-frame vector to the ring / GRID) are judged against the magnitude of
        # their operands: a drone at x = 29 carries half an ulp(29) = 1e-6 of rounding in x whatever the
        # difference comes out as.  Every other column: ABS_TOL + REL_TOL * |reference value|.
        scale = np.abs(robs).astype(np.float64)
        pos = np.abs(rag[:, 0:3]).astype(np.float64)
        scale[:, 23:26] = np.maximum(scale[:, 23:26], pos)
        scale[:, 26:29] = np.maximum(scale[:, 26:29], pos / grid)
        scale[:, 32:35] = np.maximum(scale[:, 32:35], pos)
        scale[:, 35:38] = np.maximum(scale[:, 35:38], pos.max(axis=1, keepdims=True) / grid)
        tol = _tol_err(obs, robs, scale)
        ok = tol <= 1.0
        bad_rows = ~ok.all(axis=1)
        only_neighbour = ok[:, :32].all(axis=1) & ok[:, 35:].all(axis=1)  # obs[32:35] = clamp(nearest - pos)
        stats["nearest_neighbour_flips"] += int((bad_rows & only_neighbour).sum())
        good = ~(bad_rows & only_neighbour)
        stats["worst_obs_tol"] = max(stats["worst_obs_tol"], float(tol[good].max()))
        stats["worst_reward_abs"] = max(stats["worst_reward_abs"], float(np.abs(rew - rrew).max()))
        live = rterm == 0
        stats["worst_state_tol"] = max(stats["worst_state_tol"], _state_err(ag[live, :17], rag[live, :17]))
    stats["guard_replays_whole_vector"] = vec.guard_replays
    stats["config"] = f"SwarmVec({n}, {A}, max_rings={R}, math='fast'), env slices {slices}, {T} steps, oracle = CPU restatement (Philox draws)"
    stats["tolerance"] = {"rel": REL_TOL, "abs": ABS_TOL}
    _record(f"swarm_fast_per_step_A{A}", stats)
    for o in orcs:
        o.close()
    vec.close()
    assert stats["drone_steps"] >= 1_000_000
    for k in ("terminal_flips", "ring_idx_flips", "episode_length_flips", "collision_count_flips", "nearest_neighbour_flips"):
        assert stats[k] == 0, f"{k} = {stats[k]} over {stats['drone_steps']} drone-steps (integer outputs must be identical)"
    assert stats["worst_obs_tol"] <= 1.0, stats
    assert stats["worst_reward_abs"] <= 2e-5, stats
    assert stats["worst_state_tol"] <= 1.0, stats
