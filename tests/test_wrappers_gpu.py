"""The PufferLib-shaped front door on the GPU: DroneRace / DroneSwarm wrappers over the CPython
`binding` module (reference: pufferlib/ocean/drone_race/drone_race.py, drone_swarm/drone_swarm.py,
env_binding.h), with the reference's NumPy buffer contract and with zero-copy device buffers."""
import numpy as np
import pytest

from _util import action_tape, bits

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_drone_race_numpy_contract_matches_oracle(oracle):
    from drone_b200.drone_race import DroneRace
    n, seed = 1500, 5
    env = DroneRace(num_envs=n, seed=seed, math="strict", report_interval=16)
    assert env.observations.shape == (n, 29) and env.observations.dtype == np.float32
    assert env.actions.dtype == np.float32 and env.terminals.dtype == bool and env.num_agents == n
    cpu = oracle.OrcRace(n, seed=seed)
    cpu.reset(seed, mode=oracle.RESET_PHILOX)
    obs, infos = env.reset(seed)
    assert infos == [] and obs is env.observations
    assert np.array_equal(bits(obs), bits(cpu.observations))
    tape = action_tape(n, scale=1.0)
    logged = 0
    for t in range(64):
        cpu.step(tape[t % 16], mode=oracle.RESET_PHILOX)
        obs, rew, term, trunc, info = env.step(tape[t % 16])
        assert np.array_equal(bits(obs), bits(cpu.observations)), f"obs differ at step {t}"
        assert np.array_equal(bits(rew), bits(cpu.rewards)) and np.array_equal(term.astype(np.uint8), cpu.terminals)
        assert not trunc.any()
        if info:
            logged += 1
            assert set(info[0]) == {"perf", "score", "collision_rate", "oob", "timeout", "episode_return",
                                    "episode_length", "n"}  # drone_race/binding.c:13-23
    assert logged >= 1
    with pytest.raises(TypeError):
        env.reset(None)  # env_binding.h:494-497: seed must be an int
    env.close()
    cpu.close()


def test_drone_race_reference_construction_path_env_init_vectorize(oracle):
    """drone_race.py:37-51: one env_init per env on buffer slices, then vectorize(*handles)."""
    from drone_b200.drone_race import DroneRace
    n, seed = 96, 0
    a = DroneRace(num_envs=n, seed=seed, math="strict", per_env_init=True)
    b = DroneRace(num_envs=n, seed=seed, math="strict")
    a.reset(3)
    b.reset(3)
    tape = action_tape(n, scale=1.0)
    for t in range(20):
        oa = a.step(tape[t % 16])[0]
        ob = b.step(tape[t % 16])[0]
    assert np.array_equal(bits(oa), bits(ob))
    a.close()
    b.close()


def test_drone_race_device_buffers_are_zero_copy():
    from drone_b200.drone_race import DroneRace
    n = 4096
    env = DroneRace(num_envs=n, buffers="device", seed=1, report_interval=1 << 30)
    env.reset(1)
    assert env.observations.is_cuda and env.observations.shape == (n, 29)
    ptr = env.observations.data_ptr()
    acts = torch.rand((n, 4), device="cuda") * 2 - 1
    obs, rew, term, trunc, info = env.step(acts)
    assert obs.data_ptr() == ptr and obs is env.observations  # written in place, no copy
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and float(obs.abs().sum()) > 0
    cap = torch.utils.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(env.observations))
    assert cap.data_ptr() == ptr
    env.close()


def test_drone_swarm_numpy_contract_matches_oracle(oracle):
    from drone_b200.drone_swarm import DroneSwarm
    E, A, R, seed = 12, 16, 5, 7
    env = DroneSwarm(num_envs=E, num_drones=A, max_rings=R, seed=seed, math="strict", report_interval=8)
    assert env.num_agents == E * A and env.observations.shape == (E * A, 41)
    cpu = oracle.OrcSwarm(E, A, R, seed=seed)
    cpu.reset(seed, mode=oracle.RESET_PHILOX)
    obs, _ = env.reset(seed)
    assert np.array_equal(bits(obs), bits(cpu.observations))
    tape = action_tape(E * A, scale=1.0)
    keys = None
    for t in range(40):
        cpu.step(tape[t % 16], mode=oracle.RESET_PHILOX)
        obs, rew, term, trunc, info = env.step(tape[t % 16])
        assert np.array_equal(bits(obs), bits(cpu.observations)), f"obs differ at step {t}"
        assert np.array_equal(bits(rew), bits(cpu.rewards)) and np.array_equal(term.astype(np.uint8), cpu.terminals)
        if info:
            keys = set(info[0])
    assert keys == {"perf", "score", "rings_passed", "collision_rate", "oob", "episode_return", "episode_length", "n"}
    env.close()
    cpu.close()


def test_drone_swarm_per_env_init_and_device_buffers():
    from drone_b200.drone_swarm import DroneSwarm
    E, A = 6, 8
    a = DroneSwarm(num_envs=E, num_drones=A, seed=2, math="strict", per_env_init=True)
    b = DroneSwarm(num_envs=E, num_drones=A, seed=2, math="strict", buffers="device")
    a.reset(9)
    b.reset(9)
    tape = action_tape(E * A, scale=1.0)
    for t in range(10):
        oa = a.step(tape[t % 16])[0]
        ob = b.step(torch.from_numpy(tape[t % 16]).cuda())[0]
    torch.cuda.synchronize()
    assert np.array_equal(bits(oa), bits(ob.cpu().numpy()))
    a.close()
    b.close()


def test_capturable_step_in_cuda_graph():
    """vec_step is allocation-free and sync-free: K steps captured once, replayed."""
    from drone_b200.vec import RaceVec
    n, K = 8192, 8
    vec = RaceVec(n, seed=3, math="strict")
    ref = RaceVec(n, seed=3, math="strict")
    vec.reset(3)
    ref.reset(3)
    tape = torch.from_numpy(action_tape(n, scale=1.0)).cuda()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                vec.step(tape[k], stream=s)
    torch.cuda.synchronize()
    assert vec.step_count == 0  # capture does not execute
    g.replay()
    g.replay()
    for rep in range(2):
        for k in range(K):
            ref.step(tape[k])
    torch.cuda.synchronize()
    assert vec.step_count == ref.step_count == 2 * K
    assert np.array_equal(bits(vec.get_state()), bits(ref.get_state()))
    assert np.array_equal(bits(vec.observations.cpu().numpy()), bits(ref.observations.cpu().numpy()))
    vec.close()
    ref.close()


def test_host_buffer_pipeline_equals_device_path():
    """NumPy-buffer steps of large vectors are issued as a chunked H2D / kernel / D2H pipeline
    (tile sub-range launches); results, step count and vec_log must equal the one-launch path."""
    from drone_b200.drone_race import DroneRace
    from drone_b200.vec import RaceVec
    n, seed, T = 140_003, 11, 12
    env = DroneRace(num_envs=n, seed=seed, math="strict", max_moves=9, report_interval=1 << 30)
    vec = RaceVec(n, seed=seed, math="strict", max_moves=9)
    env.reset(seed)
    vec.reset(seed)
    tape = action_tape(n, scale=1.0)
    dtape = torch.from_numpy(tape).cuda()
    for t in range(T):
        obs, rew, term, trunc, info = env.step(tape[t % 16])
        vec.step(dtape[t % 16])
    torch.cuda.synchronize()
    assert np.array_equal(bits(obs), bits(vec.observations.cpu().numpy()))
    assert np.array_equal(bits(rew), bits(vec.rewards.cpu().numpy()))
    assert np.array_equal(term.astype(np.uint8), vec.terminals.cpu().numpy())
    assert np.array_equal(bits(env.actions), bits(tape[(T - 1) % 16]))  # the shared action buffer holds the actions
    from drone_b200.drone_race import binding
    assert binding.vec_log(env.c_envs) == {k: v for k, v in vec.log().items()}
    assert vec.step_count == T
    env.close()
    vec.close()


@pytest.mark.parametrize("A", [16, 64, 5])
def test_swarm_host_buffer_pipeline_equals_device_path(A):
    """NumPy-buffer steps of large swarm vectors are issued as the same chunked H2D / kernel / D2H pipeline as the
    race env's (tile sub-range launches of swarm_kernel); results, buffers and vec_log must equal the one-launch path."""
    from drone_b200.drone_swarm import DroneSwarm, binding
    from drone_b200.vec import SwarmVec
    E, R, seed, T = (131_200 + A - 1) // A + 3, 6, 9, 10
    env = DroneSwarm(num_envs=E, num_drones=A, max_rings=R, seed=seed, math="strict", report_interval=1 << 30)
    vec = SwarmVec(E, A, R, seed=seed, math="strict")
    env.reset(seed)
    vec.reset(seed)
    tape = action_tape(E * A, scale=1.0)
    dtape = torch.from_numpy(tape).cuda()
    for t in range(T):
        obs, rew, term, trunc, info = env.step(tape[t % 16])
        vec.step(dtape[t % 16])
    torch.cuda.synchronize()
    assert np.array_equal(bits(obs), bits(vec.observations.cpu().numpy()))
    assert np.array_equal(bits(rew), bits(vec.rewards.cpu().numpy()))
    assert np.array_equal(term.astype(np.uint8), vec.terminals.cpu().numpy())
    assert np.array_equal(bits(env.actions), bits(tape[(T - 1) % 16]))
    assert binding.vec_log(env.c_envs) == {k: v for k, v in vec.log().items()}
    assert vec.step_count == T
    env.close()
    vec.close()


def test_host_step_from_a_pinned_caller_array_uploads_in_place_and_leaves_clamped_actions():
    """A page-locked action array (env.pinned_actions()) is uploaded from where it is, without the staging copy
    (api.cu step_host_impl); the env's own action buffer must still hold clamp(actions, -1, 1) at return, like the
    reference's (dronelib.h:437), and the results must equal those of ordinary (copied) arrays.  The library never
    page-locks caller memory itself: arrays that are freed and re-allocated at the same address stay correct."""
    from drone_b200.drone_race import DroneRace
    n, seed = 140_003, 5
    a = DroneRace(num_envs=n, seed=seed, math="strict", report_interval=1 << 30)
    b = DroneRace(num_envs=n, seed=seed, math="strict", report_interval=1 << 30)
    a.reset(seed)
    b.reset(seed)
    rng = np.random.default_rng(3)
    pinned = a.pinned_actions()
    assert pinned.shape == (n, 4) and pinned.dtype == np.float32
    for t in range(5):
        fresh = rng.uniform(-1.6, 1.6, size=(n, 4)).astype(np.float32)
        fresh[7, 2] = np.nan
        pinned[:] = fresh
        oa, ra, ta, _, _ = a.step(pinned)          # in-place upload
        ob, rb, tb, _, _ = b.step(fresh.copy())    # a new pageable array every step (same address, most likely): copy path
        assert np.array_equal(bits(oa), bits(ob)) and np.array_equal(bits(ra), bits(rb)) and np.array_equal(ta, tb), t
        want = np.clip(fresh, -1.0, 1.0)
        assert np.array_equal(bits(a.actions), bits(want)), t
        assert np.array_equal(bits(b.actions), bits(want)), t
        assert np.array_equal(bits(pinned), bits(fresh)), t  # the caller's array is read, never written
    a.close()
    b.close()


def test_registry_make_builds_the_ini_default_envs():
    from drone_b200 import registry
    env = registry.make("puffer_drone_swarm", env_kwargs=dict(num_envs=4), seed=3)
    assert env.num_agents == 4 * 64 and env.observations.shape == (256, 41)  # drone_swarm.ini: 64 drones
    obs, _ = env.reset(3)
    env.step(np.zeros((256, 4), np.float32))
    assert np.isfinite(env.observations).all()
    env.close()
    env = registry.make("puffer_drone_race", env_kwargs=dict(num_envs=64, buffers="device"))
    env.reset(0)
    assert env.observations.is_cuda and env.num_agents == 64
    env.close()
