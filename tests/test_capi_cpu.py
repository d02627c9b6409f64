"""Without a GPU: the C-ABI library loads, exports every symbol include/mpc_b200.h declares, and
refuses to run (no CPU path) with the documented error codes.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import helpers


def _capi():
    from mpc_rl_for_avs_b200 import _capi
    return _capi


def test_library_exports_every_declared_symbol():
    capi = _capi()
    hdr = open(os.path.join(helpers.ROOT, "include", "mpc_b200.h")).read()
    declared = sorted(set(re.findall(r"MPC_API\s+[\w\s\*]+?\b(mpc_\w+)\s*\(", hdr)))
    assert len(declared) >= 14
    assert sorted(capi.EXPORTS) == declared
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_the_header():
    capi = _capi()
    assert C.sizeof(capi.MpcConfig) == 17 * 4
    assert C.sizeof(capi.MpcProblemBatch) == 12 * C.sizeof(C.c_void_p)
    assert C.sizeof(capi.MpcSolveOut) == 5 * C.sizeof(C.c_void_p)
    assert C.sizeof(capi.MpcLatchState) == 3 * C.sizeof(C.c_void_p)
    assert C.sizeof(capi.MpcCollisionOut) == 7 * C.sizeof(C.c_void_p)


def test_create_rejects_bad_arguments_and_missing_device():
    import torch
    capi = _capi()
    lib = capi.load()
    h = C.c_void_p()
    good = dict(abi_version=capi.ABI_VERSION, horizon=20, vehicles_count=9, dt=0.1, weight_speed=1, weight_control=1,
                weight_input_diff=1)
    for bad in (dict(abi_version=99), dict(horizon=1), dict(horizon=65), dict(vehicles_count=0), dict(vehicles_count=18), dict(dt=0.0), dict(dt=0.15), dict(n_starts=9), dict(n_starts=-1)):
        cfg = capi.MpcConfig(**{**good, **bad})
        assert lib.mpc_create(C.byref(cfg), 0, 16, C.byref(h)) == capi.ERR_BAD_ARG
        assert lib.mpc_last_error(None)
    cfg = capi.MpcConfig(**good)
    assert lib.mpc_create(C.byref(cfg), 0, 0, C.byref(h)) == capi.ERR_BAD_ARG
    if not torch.cuda.is_available():
        rc = lib.mpc_create(C.byref(cfg), 0, 16, C.byref(h))
        assert rc == capi.ERR_NO_DEVICE and b"no CPU path" in lib.mpc_last_error(None)
        assert not h.value


def test_python_surface_fails_loudly_without_cuda():
    import torch
    import mpc_rl_for_avs_b200 as pkg
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        pkg.BatchedPureMPC({"horizon": 20}, vehicles_count=9, max_batch=4)

    class Env:
        config = {"simulation_frequency": 30, "policy_frequency": 10, "observation": {"vehicles_count": 9}}
    env = Env(); env.unwrapped = env
    with pytest.raises(RuntimeError, match="no CPU path"):
        pkg.PureMPC_Agent(env, {"horizon": 16, "render": False, "weight_speed": 1, "weight_control": 1, "weight_input_diff": 1})


def test_product_never_imports_the_oracle():
    """The oracle and the host harness are test infrastructure: nothing under the package may touch them."""
    pkg_dir = os.path.join(helpers.ROOT, "mpc-rl_for_avs_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "mpc_oracle" not in txt and "hostsim" not in txt.replace("tests/hostsim", "") or f == "mpc_core.cuh", f
                assert "import oracle" not in txt and "scipy" not in txt, f


def test_scenarios_are_seeded_and_well_formed():
    import numpy as np
    import mpc_rl_for_avs_b200 as pkg
    a = pkg.make_scenarios(64, 8, seed=3)
    b = pkg.make_scenarios(64, 8, seed=3)
    c = pkg.make_scenarios(64, 8, seed=4)
    assert all((x == y).all() for x, y in zip(a, b)) and not (a[0] == c[0]).all()
    obs = a[0].numpy()
    assert obs.shape == (64, 9, 8) and obs.dtype == np.float32 and (obs[:, :, 0] == 1).all()
    d = np.hypot(obs[:, 1:, 1] - obs[:, :1, 1], obs[:, 1:, 2] - obs[:, :1, 2])
    assert d.min() >= 5.0 - 1e-4
    assert np.abs(obs[:, 0, 5]).max() <= np.pi + 1e-6
