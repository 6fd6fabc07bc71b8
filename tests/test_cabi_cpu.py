"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/b200drone.h declares, the CPython `binding` modules export the reference's
method table (env_binding.h:644-662) and raise the reference's exception types for the
same mistakes (pinned upstream by tests/test_env_binding.py), and the product path fails
loudly -- never falls back to a CPU implementation -- when there is no CUDA device.
No compute call is made here.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200drone.h")

REFERENCE_METHODS = ["env_init", "env_reset", "env_step", "env_render", "env_close", "env_get", "env_put",
                     "vectorize", "vec_init", "vec_reset", "vec_step", "vec_log", "vec_render", "vec_close", "shared"]


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2d_[a-z0-9_]+)\s*\(", src)))


def _have_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    from drone_b200 import capi
    lib = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200drone.h but not exported"
    assert set(capi.SYMBOLS) == set(declared), "drone_b200/capi.py and include/b200drone.h disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (b2d_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert lib.b2d_version() == 1


def test_library_is_sm100a_only_and_links_no_oracle():
    from drone_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    assert not re.search(r"sm_(?!100a)\d+", out), "only sm_100a code is shipped"
    ldd = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "libref" not in ldd


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "drone_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"
                assert "liboracle" not in src and "libref_" not in src


@pytest.mark.parametrize("flavour", ["drone_race", "drone_swarm"])
def test_binding_module_method_table(flavour):
    import importlib
    binding = importlib.import_module(f"drone_b200.{flavour}.binding")
    for m in REFERENCE_METHODS:
        assert callable(getattr(binding, m)), f"binding.{m} missing (env_binding.h:644-662)"


def _bufs(n=4, obs=29, atn_dtype=np.float32):
    return (np.zeros((n, obs), np.float32), np.zeros((n, 4), atn_dtype), np.zeros(n, np.float32),
            np.zeros(n, bool), np.zeros(n, bool))


def test_binding_argument_errors_match_reference_convention():
    """TypeError for arity / kwargs / non-int handles, ValueError for layout
    (reference: tests/test_env_binding.py:63-113, env_binding.h:51-135,615-622)."""
    from drone_b200.drone_race import binding
    o, a, r, t, tr = _bufs()
    with pytest.raises(TypeError):
        binding.env_init()
    with pytest.raises(TypeError):
        binding.env_init(o[:1], a[:1], r[:1], t[:1], tr[:1])  # seed missing
    with pytest.raises(TypeError):
        binding.env_init(o[:1], a[:1], r[:1], t[:1], tr[:1], 0)  # required kwargs missing
    with pytest.raises(TypeError):
        binding.env_init(o[:1], a[:1], r[:1], t[:1], tr[:1], 0, max_rings=10)  # max_moves missing
    with pytest.raises(TypeError):
        binding.env_init(o[:1], a[:1], r[:1], t[:1], tr[:1], "seed", max_rings=10, max_moves=1000)
    with pytest.raises(TypeError):
        binding.env_init([0.0] * 29, a[:1], r[:1], t[:1], tr[:1], 0, max_rings=10, max_moves=1000)
    with pytest.raises(ValueError):
        binding.env_init(o[:1], np.zeros((1, 4), np.float64), r[:1], t[:1], tr[:1], 0, max_rings=10, max_moves=1000)
    with pytest.raises(ValueError):
        binding.env_init(o[:, ::2][:1], a[:1], r[:1], t[:1], tr[:1], 0, max_rings=10, max_moves=1000)
    with pytest.raises(ValueError):
        binding.env_init(o[:1], a[:1], np.zeros((1, 1), np.float32), t[:1], tr[:1], 0, max_rings=10, max_moves=1000)
    with pytest.raises(TypeError):
        binding.vec_init(o, a, r, t, tr, 4)  # seed missing
    with pytest.raises(TypeError):
        binding.vec_init(o, a, r, t, tr, 0, 0, max_rings=10, max_moves=1000)  # num_envs must be > 0
    with pytest.raises(ValueError):
        binding.vec_init(np.zeros(29, np.float32), a, r, t, tr, 4, 0, max_rings=10, max_moves=1000)
    with pytest.raises(TypeError):
        binding.vectorize()
    with pytest.raises(TypeError):
        binding.vectorize([1, 2])
    with pytest.raises(TypeError):
        binding.vec_step()
    with pytest.raises(TypeError):
        binding.vec_step("handle")
    with pytest.raises(TypeError):
        binding.vec_reset(12345)  # arity
    # env_init does no device work: a handle can be made and closed on any machine
    h = binding.env_init(o[:1], a[:1], r[:1], t[:1], tr[:1], 0, max_rings=10, max_moves=1000)
    assert isinstance(h, int) and h != 0
    binding.env_close(h)


def test_swarm_binding_requires_its_kwargs():
    from drone_b200.drone_swarm import binding
    o, a, r, t, tr = _bufs(8, 41)
    with pytest.raises(TypeError):
        binding.env_init(o, a, r, t, tr, 0, max_rings=10)  # num_agents missing (DS/binding.c:7-8)
    h = binding.env_init(o, a, r, t, tr, 0, num_agents=8, max_rings=10)
    binding.env_close(h)


def test_no_cpu_fallback_without_a_device():
    """On a machine without CUDA every compute entry point must fail loudly."""
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    from drone_b200 import capi
    from drone_b200.drone_race import binding
    lib = capi.lib()
    cfg = capi.RaceCfg(16, 10, 1000, 0, 0, 0, capi.MATH_FAST, 0)
    h = C.c_void_p()
    rc = lib.b2d_race_create(C.byref(h), C.byref(cfg), None)
    assert rc == capi.B2D_ECUDA and not h.value
    assert lib.b2d_last_error()
    o, a, r, t, tr = _bufs()
    with pytest.raises(RuntimeError):
        binding.vec_init(o, a, r, t, tr, 4, 0, max_rings=10, max_moves=1000)
    from drone_b200.vec import RaceVec
    with pytest.raises(RuntimeError):
        RaceVec(16)
    from drone_b200.drone_race import DroneRace
    with pytest.raises(RuntimeError):
        DroneRace(num_envs=4)


def test_create_argument_validation_happens_before_cuda():
    from drone_b200 import capi
    lib = capi.lib()
    h = C.c_void_p()
    for cfg in (capi.RaceCfg(0, 10, 1000, 0, 0, 0, 0, 0), capi.RaceCfg(8, 0, 1000, 0, 0, 0, 0, 0),
                capi.RaceCfg(8, 10, 0, 0, 0, 0, 0, 0), capi.RaceCfg(8, 10, 10, 0, 0, 0, 7, 0)):
        assert lib.b2d_race_create(C.byref(h), C.byref(cfg), None) == capi.B2D_EINVAL
    for scfg in (capi.SwarmCfg(4, 0, 5, 0, 0, 0, 0, 0), capi.SwarmCfg(4, 129, 5, 0, 0, 0, 0, 0),
                 capi.SwarmCfg(0, 8, 5, 0, 0, 0, 0, 0), capi.SwarmCfg(4, 8, 0, 0, 0, 0, 0, 0)):
        assert lib.b2d_swarm_create(C.byref(h), C.byref(scfg), None) == capi.B2D_EINVAL
    assert lib.b2d_race_create(None, None, None) == capi.B2D_EINVAL
    assert lib.b2d_vec_step(None, None) == capi.B2D_EINVAL
    assert lib.b2d_vec_close(None) == capi.B2D_EINVAL
    with pytest.raises(ValueError):
        capi.check(capi.B2D_EINVAL)
    with pytest.raises(MemoryError):
        capi.check(capi.B2D_ENOMEM)
