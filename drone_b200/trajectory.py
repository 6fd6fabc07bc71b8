"""Env-state checkpoints and trajectory dumps for device-resident envs.

The reference's viewer (c_render, pufferlib/ocean/drone_race/drone_race.h:331-462) draws from the env's host
structs; envs of this package live in HBM.  Two bridges:

  export_ref(vec, env_id)   one env as the reference's own `Drone` / `Ring` structs (capi.RefDrone / RefRing,
                            byte-compatible with dronelib.h:161-166,191-247) -- what a viewer or a debugger
                            written against the reference consumes;
  record(vec, steps, ...)   a trajectory [T + 1, n, blob] of full state blobs (b2d_get_state layout, the same
                            blob b2d_put_state restores) plus the per-step actions / rewards / terminals, saved as
                            .npz: replayable on any build, diffable against the oracle.
"""
import ctypes as C

import numpy as np

from . import capi


def blob_to_ref(vec, blob):
    """One state blob -> (drones, rings, tick, aux): ctypes arrays of RefDrone / RefRing; aux = ring_idx (race) or task (swarm)."""
    blob = np.ascontiguousarray(blob, np.float32)
    R = vec.max_rings
    rings = (capi.RefRing * R)()
    tick, aux = C.c_int(), C.c_int()
    fp = blob.ctypes.data_as(C.POINTER(C.c_float))
    if vec.obs_dim == 29:
        drones = (capi.RefDrone * 1)()
        ret = C.c_float()
        capi.check(capi.lib().b2d_race_blob_to_ref(fp, R, drones, rings, C.byref(tick), C.byref(aux), C.byref(ret)))
    else:
        drones = (capi.RefDrone * vec.num_drones)()
        capi.check(capi.lib().b2d_swarm_blob_to_ref(fp, vec.num_drones, R, drones, rings, C.byref(tick), C.byref(aux)))
    return drones, rings, tick.value, aux.value


def export_ref(vec, env_id):
    """The live state of env `env_id` as reference structs (synchronous)."""
    R = vec.max_rings
    rings = (capi.RefRing * R)()
    drones = (capi.RefDrone * (1 if vec.obs_dim == 29 else vec.num_drones))()
    tick, aux, ret = C.c_int(), C.c_int(), C.c_float()
    capi.check(capi.lib().b2d_export_ref(vec.h, int(env_id), drones, rings, C.byref(tick), C.byref(aux), C.byref(ret)))
    return drones, rings, tick.value, aux.value


def record(vec, steps, env_ids, actions, path=None):
    """Step `vec` `steps` times with actions(t) -> CUDA tensor [num_agents, 4] and record the envs in `env_ids`:
    states [steps + 1, n, blob], and per step the rows of those envs in actions / rewards / terminals.
    Returns the dict (and writes it to `path` as .npz when given)."""
    import torch
    env_ids = [int(i) for i in env_ids]
    per = vec.num_agents // vec.num_envs
    rows = np.concatenate([np.arange(e * per, (e + 1) * per) for e in env_ids])
    trow = torch.as_tensor(rows, device=vec.device)
    states = [vec.get_state(env_ids)]
    acts, rews, terms = [], [], []
    for t in range(steps):
        a = actions(t)
        vec.step(a)
        acts.append(a[trow].cpu().numpy())
        rews.append(vec.rewards[trow].cpu().numpy())
        terms.append(vec.terminals[trow].cpu().numpy())
        states.append(vec.get_state(env_ids))
    out = dict(states=np.stack(states), actions=np.stack(acts), rewards=np.stack(rews), terminals=np.stack(terms),
               env_ids=np.asarray(env_ids), max_rings=np.asarray(vec.max_rings), agents_per_env=np.asarray(per))
    if path is not None:
        np.savez_compressed(path, **out)
    return out
