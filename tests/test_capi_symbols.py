"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/dockauv.h declares, agrees on
the parameter-block size, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dockauv.h")).read()
    return sorted(set(re.findall(r"DOCKAUV_API[^;(]*?\b(dockauv_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from gym_dockauv_b200 import _capi
    lib = _capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dockauv.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared, "python binding table and header disagree"


def test_struct_sizes_and_abi():
    from gym_dockauv_b200 import _capi
    from gym_dockauv_b200.params import ABI_VERSION, DockauvParams
    lib = _capi.load()
    assert lib.dockauv_abi_version() == ABI_VERSION
    assert lib.dockauv_sizeof_params() == C.sizeof(DockauvParams)


def test_argument_validation_without_gpu():
    from gym_dockauv_b200 import _capi
    from gym_dockauv_b200.config import BASE_CONFIG
    from gym_dockauv_b200.params import pack_params
    lib = _capi.load()
    p, meta = pack_params(BASE_CONFIG, "ObstaclesDocking3d")
    assert lib.dockauv_n_obs(C.byref(p)) == meta["n_obs"] == 36
    h = C.c_void_p()
    p.abi_version = 99
    assert lib.dockauv_create(C.byref(p), 16, 0, C.byref(h)) == -1 and b"ABI" in lib.dockauv_last_error()
    from gym_dockauv_b200.params import ABI_VERSION
    p.abi_version = ABI_VERSION
    assert lib.dockauv_create(C.byref(p), 0, 0, C.byref(h)) == -1
    p.n_rays = 62
    assert lib.dockauv_create(C.byref(p), 16, 0, C.byref(h)) == -1 and b"radar" in lib.dockauv_last_error()
    # every value that would silently produce inf / NaN observations or read uninitialised reset draws is refused
    for field, bad, word in (("n_synthetic_spheres", -1, b"n_synthetic_spheres"), ("n_synthetic_spheres", 1, b"n_synthetic_spheres"),
                             ("h", 0.0, b"t_step_size"), ("max_timesteps", 0, b"max_timesteps"), ("u_max", 0.0, b"u_max"),
                             ("r_max", -1.0, b"u_max"), ("max_attitude", 0.0, b"max_attitude"),
                             ("max_dist_from_goal", 0.0, b"max_dist_from_goal"), ("dist_goal_reached_tol", 0.0, b"dist_goal_reached_tol"),
                             ("radar_max_dist", 0.0, b"max_dist"), ("layout", 3, b"layout"), ("reward_set", 3, b"reward_set")):
        q, _ = pack_params(BASE_CONFIG, "ObstaclesDocking3d")
        setattr(q, field, bad)
        assert lib.dockauv_create(C.byref(q), 16, 0, C.byref(h)) == -1, field
        assert word in lib.dockauv_last_error(), (field, lib.dockauv_last_error())
    assert lib.dockauv_step(None, None, 0, None, None, None, 0, None) == -1
    assert lib.dockauv_destroy(None) == 0
    # the caller-side entry points validate their arguments before touching the device
    assert lib.dockauv_rollout(None, None, 0, 4, None, 1, 1, None) == -1
    assert lib.dockauv_gae(None, 0, None, None, None, 4, 16, 0.99, 0.95, None, None, None) == -1
    assert lib.dockauv_fold_stats(None, None) == -1
    n = C.c_int()
    ms = (C.c_float * 4)()
    assert lib.dockauv_last_step_launch_ms(None, ms, 4, C.byref(n)) == -1
    assert lib.dockauv_step_host(None, None, 0, None, None, None, None, 1, None) == -1
    assert lib.dockauv_refresh_obstacles(None, None) == -1
    assert lib.dockauv_rollout_captures(None, None) == -1
    assert lib.dockauv_last_list_counts(None, None, None, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gym_dockauv_b200 import _capi, envs
    from gym_dockauv_b200.config import BASE_CONFIG
    from gym_dockauv_b200.params import pack_params
    lib = _capi.load()
    p, _ = pack_params(BASE_CONFIG, "SimpleDocking3d")
    h = C.c_void_p()
    assert lib.dockauv_create(C.byref(p), 16, 0, C.byref(h)) == -2      # DOCKAUV_ECUDA
    assert b"no CPU fallback" in lib.dockauv_last_error()
    with pytest.raises(_capi.DockauvError):
        envs.SimpleDocking3d(BASE_CONFIG, num_envs=4)
    with pytest.raises(_capi.DockauvError):
        envs.SimpleDocking3d(BASE_CONFIG, num_envs=4, device="cpu")
    with pytest.raises(KeyError):
        envs.make_gym("Nope-v0", BASE_CONFIG)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gym_dockauv_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gym_dockauv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
