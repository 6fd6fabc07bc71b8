"""Parity at BASELINE.json's FULL sizes (configs[1]: 1,048,576 race envs; configs[2]: 65,536 swarm
envs x 64 drones) through size-independent properties.

The oracle cannot step a million envs in seconds, but the device reset stream makes every env a
pure function of (seed, global env id, its own actions) -- so a slice [a, b) of the full-size
vector must equal, bit for bit, what the CPU restatement computes for (b - a) envs created with
env_id_base = a.  Slices sit at the start, at an unaligned offset in the middle, and across the
ragged tail.  On top of that: launch-mode invariance (tape == separate steps), log conservation
(episodes counted by vec_log == terminals raised, the checksum of checksums), state invariants
(unit quaternions, clamps, tick and ring bounds) in the fast build the bench runs.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

RACE_N = 1 << 20          # BASELINE.json configs[1]
SWARM_ENVS, SWARM_A = 1 << 16, 64  # BASELINE.json configs[2]


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _race_tape(n, steps=16, seed=1234):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.rand((steps, n, 4), generator=g) * 2.0 - 1.0  # bench.py's tape recipe


def test_race_full_size_slices_equal_the_oracle(oracle):
    from drone_b200.vec import RaceVec
    n, seed, T1, T2 = RACE_N, 0, 24, 40
    tape = _race_tape(n)
    dtape = tape.cuda()
    htape = tape.numpy()
    vec = RaceVec(n, max_rings=10, max_moves=1000, math="strict", seed=seed)
    vec.reset(seed)
    slices = [(0, 4096), (517_123, 517_123 + 4099), (n - 4097, n)]
    cpus = []
    for a, b in slices:
        cpu = oracle.OrcRace(b - a, max_rings=10, max_moves=1000, seed=seed, env_id_base=a)
        cpu.reset(seed, mode=oracle.RESET_PHILOX)
        cpus.append(cpu)
    obs0 = vec.observations.cpu().numpy()
    for (a, b), cpu in zip(slices, cpus):
        assert np.array_equal(_bits(obs0[a:b]), _bits(cpu.observations)), f"reset obs differ in [{a},{b})"
    nterm = torch.zeros((), dtype=torch.int64, device="cuda")
    done = 0
    for T in (T1, T2):
        for t in range(done, T):
            vec.step(dtape[t % 16])
            nterm += vec.terminals.sum()
            for (a, b), cpu in zip(slices, cpus):
                cpu.step(htape[t % 16, a:b], mode=oracle.RESET_PHILOX)
        done = T
        obs, rew, term = vec.observations.cpu().numpy(), vec.rewards.cpu().numpy(), vec.terminals.cpu().numpy()
        for (a, b), cpu in zip(slices, cpus):
            tag = f"step {T}, envs [{a},{b})"
            assert np.array_equal(term[a:b], cpu.terminals), tag
            assert np.array_equal(_bits(rew[a:b]), _bits(cpu.rewards)), tag
            assert np.array_equal(_bits(obs[a:b]), _bits(cpu.observations)), tag
            assert np.array_equal(_bits(vec.get_state(range(a, b))), _bits(cpu.get_state())), tag
    # checksum of checksums: every terminal raised is one episode in vec_log, nothing else is
    log = vec.log()
    assert log["n"] == float(nterm.item()) and log["n"] > 0
    assert vec.step_count == T2
    for cpu in cpus:
        cpu.close()
    vec.close()


def test_race_full_size_tape_equals_steps_and_invariants():
    """bench.py's launch mode (b2d_vec_step_tape, overlapped launches) at bench.py's size and math:
    identical to separate vec_steps, and the state obeys the env's invariants."""
    from drone_b200.vec import RaceVec
    n, seed, T = RACE_N, 0, 64
    dtape = _race_tape(n).cuda()
    a = RaceVec(n, math="fast", seed=seed)
    b = RaceVec(n, math="fast", seed=seed)
    a.reset(seed)
    b.reset(seed)
    for t in range(T):
        a.step(dtape[t % 16])
    b.step_tape(dtape, 0, T)
    torch.cuda.synchronize()
    assert torch.equal(a.observations.view(torch.int32), b.observations.view(torch.int32))
    assert torch.equal(a.rewards.view(torch.int32), b.rewards.view(torch.int32))
    assert torch.equal(a.terminals, b.terminals)
    ids = list(range(0, n, 257))  # 4081 envs across the whole vector
    sa, sb = a.get_state(ids), b.get_state(ids)
    assert np.array_equal(_bits(sa), _bits(sb))
    la, lb = a.log(), b.log()
    assert la == lb and la["n"] > n  # mean episode is ~40 steps: every env has finished at least once
    # invariants (blob layout: pos3 vel3 quat4 omega3 rpm4 | 13 params | tick ring_idx ep_return | rings)
    q = sa[:, 6:10].astype(np.float64)
    assert np.abs(np.sqrt((q * q).sum(axis=1)) - 1.0).max() < 1e-5
    assert np.abs(sa[:, 0:3]).max() <= 10.0 + 1e-6          # a live drone is inside the arena
    assert np.abs(sa[:, 3:6]).max() <= 50.0 and np.abs(sa[:, 10:13]).max() <= 50.0
    assert sa[:, 30].min() >= 0 and sa[:, 30].max() <= 1000  # tick
    assert sa[:, 31].min() >= 0 and sa[:, 31].max() <= 10    # ring_idx
    obs = b.observations
    assert torch.isfinite(obs).all()
    assert float(obs[:, 21:25].abs().max()) <= 1.0 + 1e-5    # quaternion block of the observation
    assert set(torch.unique(b.terminals).tolist()) <= {0, 1}
    assert int(b.truncations.sum()) == 0                     # never written (EB:136)
    a.close()
    b.close()


def test_swarm_full_size_slices_equal_the_oracle(oracle):
    from drone_b200.vec import SwarmVec
    n, A, R, seed, T = SWARM_ENVS, SWARM_A, 10, 3, 24
    rows = n * A
    g = torch.Generator(device="cpu").manual_seed(99)
    tape = torch.rand((4, rows, 4), generator=g) * 2.0 - 1.0
    dtape, htape = tape.cuda(), tape.numpy()
    vec = SwarmVec(n, A, R, math="strict", seed=seed)
    vec.reset(seed)
    slices = [(0, 48), (30_011, 30_011 + 37), (n - 33, n)]
    orcs = []
    for a, b in slices:
        o = oracle.OrcSwarm(b - a, A, R, seed=seed, env_id_base=a)
        o.reset(seed, mode=oracle.RESET_PHILOX)
        orcs.append(o)
    obs0 = vec.observations.cpu().numpy()
    for (a, b), o in zip(slices, orcs):
        assert np.array_equal(_bits(obs0[a * A:b * A]), _bits(o.observations)), f"reset obs differ in envs [{a},{b})"
    nterm = torch.zeros((), dtype=torch.int64, device="cuda")
    for t in range(T):
        vec.step(dtape[t % 4])
        nterm += vec.terminals.sum()
        for (a, b), o in zip(slices, orcs):
            o.step(htape[t % 4, a * A:b * A], mode=oracle.RESET_PHILOX)
        if t % 6 == 5 or t == T - 1:
            obs, rew, term = vec.observations.cpu().numpy(), vec.rewards.cpu().numpy(), vec.terminals.cpu().numpy()
            for (a, b), o in zip(slices, orcs):
                tag = f"step {t}, envs [{a},{b})"
                assert np.array_equal(term[a * A:b * A], o.terminals), tag
                assert np.array_equal(_bits(rew[a * A:b * A]), _bits(o.rewards)), tag
                assert np.array_equal(_bits(obs[a * A:b * A]), _bits(o.observations)), tag
    for (a, b), o in zip(slices, orcs):
        env, ag = vec.split_state(vec.get_state(range(a, b)))
        oenv, oag = o.get_state()
        assert np.array_equal(_bits(env), _bits(oenv)) and np.array_equal(_bits(ag), _bits(oag))
        o.close()
    log = vec.log()
    assert log["n"] == float(nterm.item()) and log["n"] > 0
    vec.close()


@pytest.mark.parametrize("A", [16, 64])
def test_swarm_overlapped_launches_equal_serialised_launches(A):
    """Step launches that follow each other on one stream overlap at their edges (PDL + per-CTA completion
    flags, swarm_kernel); launches that alternate between two streams with a full synchronisation in
    between are plain.  Same vector, same actions: every output word, the state and the log must agree, in
    the fast build the bench runs, at BASELINE.json's configs[2] size."""
    from drone_b200.vec import SwarmVec
    n, R, seed, T = SWARM_ENVS, 10, 5, 40
    g = torch.Generator(device="cpu").manual_seed(17)
    dtape = (torch.rand((4, n * A, 4), generator=g) * 2.0 - 1.0).cuda()
    a = SwarmVec(n, A, R, math="fast", seed=seed)
    b = SwarmVec(n, A, R, math="fast", seed=seed)
    a.reset(seed)
    b.reset(seed)
    torch.cuda.synchronize()
    for t in range(T):          # back to back on the current stream: launches 2..T overlap their predecessor
        a.step(dtape[t % 4])
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    for t in range(T):          # a different stream every launch, drained each time: never overlapped
        b.step(dtape[t % 4], stream=s[t & 1])
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    for name in ("observations", "rewards", "terminals"):
        x, y = getattr(a, name), getattr(b, name)
        assert torch.equal(x.view(torch.uint8), y.view(torch.uint8)), name
    probe = list(range(0, 64)) + list(range(n - 64, n))
    assert np.array_equal(_bits(a.get_state(probe)), _bits(b.get_state(probe)))
    assert a.log() == b.log()
    assert a.step_count == b.step_count == T
    a.close()
    b.close()
