"""Render / checkpoint bridge (include/b200drone.h b2d_*_blob_to_ref): a state blob in the b2d_get_state
layout becomes the reference's own `Drone` / `Ring` structs, byte-compatible with dronelib.h:161-166,191-247.
Round trip: blobs recorded from the UNMODIFIED reference (tests/golden/*.npz final states) -> our structs ->
the reference's compute_observations called on those structs (oracle/_ref shims) == the observation the
reference computes from its own state, bit for bit.  The conversion is pure host code: no device needed."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from _util import GOLDEN_DIR, bits, load_golden

from drone_b200 import capi


def _ref_lib(path):
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not present")
    return C.CDLL(path)


def test_struct_sizes_match_the_reference(oracle):
    assert C.sizeof(capi.RefDrone) == 208 and C.sizeof(capi.RefRing) == 44  # probe values, SURVEY Appendix A
    L = _ref_lib(oracle.REF_RACE_SO)
    assert L.refrace_sizeof_drone() == C.sizeof(capi.RefDrone)
    assert L.refrace_sizeof_ring() == C.sizeof(capi.RefRing)
    S = _ref_lib(oracle.REF_SWARM_SO)
    assert S.refswarm_sizeof_drone() == C.sizeof(capi.RefDrone)
    assert S.refswarm_sizeof_ring() == C.sizeof(capi.RefRing)


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "race_*.npz"))))
def test_race_blob_through_reference_structs_gives_the_reference_observation(oracle, name):
    L = _ref_lib(oracle.REF_RACE_SO)
    g = load_golden(name)
    n, T, seed, R, max_moves = (int(x) for x in g["meta"])
    blobs = g["final_state"]
    ref = oracle.RefRace(n, max_rings=R, max_moves=max_moves)
    ref.put_state(blobs)
    ref.observe()
    fp = C.POINTER(C.c_float)
    for i in range(n):
        blob = np.ascontiguousarray(blobs[i], np.float32)
        drone, rings = (capi.RefDrone * 1)(), (capi.RefRing * R)()
        tick, ring_idx, ret = C.c_int(), C.c_int(), C.c_float()
        capi.check(capi.lib().b2d_race_blob_to_ref(blob.ctypes.data_as(fp), R, drone, rings, C.byref(tick), C.byref(ring_idx), C.byref(ret)))
        assert tick.value == int(blob[30]) and ring_idx.value == int(blob[31]) and ret.value == blob[32]
        obs = np.zeros(29, np.float32)
        L.refrace_observe_structs(C.byref(drone[0]), rings, R, ring_idx.value, obs.ctypes.data_as(fp))
        assert np.array_equal(bits(obs), bits(ref.observations[i])), f"env {i}"
        # what the viewer reads: unit orientation that maps +z onto the ring normal, radius 2
        for r in range(R):
            q = np.array(rings[r].orientation, np.float64)
            nrm = np.array(rings[r].normal, np.float64)
            assert abs(np.linalg.norm(q) - 1.0) < 1e-6 and rings[r].radius == 2.0
            w, x, y, z = q
            zrot = np.array([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)])
            assert np.allclose(zrot, nrm, atol=2e-6)
    ref.close()


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "swarm_*.npz"))))
def test_swarm_blob_through_reference_structs_gives_the_reference_observation(oracle, name):
    S = _ref_lib(oracle.REF_SWARM_SO)
    g = load_golden(name)
    n, A, T, seed, R = (int(x) for x in g["meta"])
    env, ag = g["final_env"], g["final_agents"]
    ref = oracle.RefSwarm(n, A, R)
    ref.put_state(env, ag)
    ref.observe()
    fp = C.POINTER(C.c_float)
    for e in range(n):
        blob = np.ascontiguousarray(np.concatenate([ag[e].reshape(-1), env[e]]), np.float32)
        drones, rings = (capi.RefDrone * A)(), (capi.RefRing * R)()
        tick, task = C.c_int(), C.c_int()
        capi.check(capi.lib().b2d_swarm_blob_to_ref(blob.ctypes.data_as(fp), A, R, drones, rings, C.byref(tick), C.byref(task)))
        assert tick.value == int(env[e, 0]) and task.value == int(env[e, 1])
        obs = np.zeros((A, 41), np.float32)
        S.refswarm_observe_structs(drones, A, rings, R, task.value, obs.ctypes.data_as(fp))
        assert np.array_equal(bits(obs), bits(ref.observations[e * A:(e + 1) * A])), f"env {e}"
    ref.close()


def test_bad_arguments_are_reported():
    with pytest.raises(ValueError):
        capi.check(capi.lib().b2d_race_blob_to_ref(None, 10, None, None, None, None, None))
