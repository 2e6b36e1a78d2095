"""The C-ABI library loads and exports every symbol include/mgym.h declares (no compute: no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from modurl_gym_b200 import build as b

    b.build()
    import modurl_gym_b200

    return modurl_gym_b200.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mgym.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgym_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from modurl_gym_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 28
    for name in names:
        assert hasattr(lib, name), f"libmgym.so does not export {name}"
    assert set(names) == set(_lib.SYMBOLS), "the ctypes table and include/mgym.h disagree"


def test_metadata(lib):
    assert lib.mgym_abi_version() == 1
    assert [lib.mgym_state_dim(k) for k in range(5)] == [4, 2, 2, 2, 4]
    assert [lib.mgym_obs_dim(k) for k in range(5)] == [4, 2, 2, 3, 6]
    assert [lib.mgym_action_is_continuous(k) for k in range(5)] == [0, 0, 1, 1, 0]
    assert [lib.mgym_num_actions(k) for k in range(5)] == [2, 3, 0, 0, 3]
    assert lib.mgym_state_dim(9) < 0
    assert lib.mgym_kind_name(0) == b"CartPole-v1" and lib.mgym_kind_name(1) == b"MountainCar-v0"


def test_spaces_match_reference_constructors(lib):
    """cartpole.rs:58-69 and mountain_car.rs:42-48."""
    from modurl_gym_b200 import spaces

    cp = spaces.observation_space(0)
    thr = np.array([0x3e567750], dtype=np.uint32).view(np.float32)[0]
    assert cp.high.tolist() == [np.float32(2.4) * 2, np.inf, thr * np.float32(2), np.inf]
    assert cp.low.tolist() == [-x for x in cp.high.tolist()]
    mc = spaces.observation_space(1)
    assert mc.low.tolist() == [np.float32(-1.2), np.float32(-0.07)] and mc.high.tolist() == [np.float32(0.6), np.float32(0.07)]
    a0, a1 = spaces.action_space(0), spaces.action_space(1)
    assert (a0.n, a1.n) == (2, 3)
    # Discrete::contains: rank-0 unsigned in range (cartpole.rs:392-403 rejects a rank-1 tensor, mountain_car.rs:374-385 value 3)
    assert a0.contains(np.uint32(1)) and not a0.contains(np.array([1], dtype=np.uint32)) and not a1.contains(np.uint32(3))


def test_config_defaults(lib):
    from modurl_gym_b200 import _lib

    want_limit = [0, 0, 999, 200, 500]
    for k in range(5):
        cfg = _lib.Config()
        assert lib.mgym_config_default(k, C.byref(cfg)) == 0
        assert cfg.struct_size == C.sizeof(_lib.Config)
        assert (cfg.auto_reset, cfg.sutton_barto_reward, cfg.is_euler, cfg.goal_velocity) == (1, 0, 1, 0.0)
        assert cfg.max_episode_steps == want_limit[k]


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.mgym_create(0, 128, 0, 1, None, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.mgym_last_error()
    import modurl_gym_b200

    with pytest.raises(Exception):
        modurl_gym_b200.GpuVecEnv("CartPole-v1", 128)


def test_bad_arguments(lib):
    h = C.c_void_p()
    assert lib.mgym_create(7, 128, 0, 1, None, C.byref(h)) == -1
    assert lib.mgym_create(0, 0, 0, 1, None, C.byref(h)) == -1
    assert lib.mgym_step(None, None, None, None, None, None, None) == -1
    assert lib.mgym_destroy(None) == 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under modurl_gym_b200/ may reference it."""
    pkg = os.path.join(ROOT, "modurl_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "mgym_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_rust_sys_declares_every_entry_point():
    """bindings/rust/src/sys.rs (uncompiled skeleton) must at least name every function of include/mgym.h."""
    text = open(os.path.join(ROOT, "bindings", "rust", "src", "sys.rs")).read()
    declared = set(re.findall(r"pub fn (mgym_[a-z_0-9]+)\(", text))
    assert declared == set(declared_symbols())


def test_rust_sources_are_self_consistent():
    """The Rust crate cannot be compiled here (no cargo/rustc), so at least: every sys:: function its wrappers call is
    declared in sys.rs, and every GpuVecEnv method the scalar Gym adapter calls exists in lib.rs."""
    src = os.path.join(ROOT, "bindings", "rust", "src")
    sys_rs, lib_rs, scalar_rs = (open(os.path.join(src, f)).read() for f in ("sys.rs", "lib.rs", "scalar.rs"))
    declared = set(re.findall(r"pub fn (mgym_[a-z_0-9]+)\(", sys_rs))
    used = set(re.findall(r"sys::(mgym_[a-z_0-9]+)\(", lib_rs + scalar_rs))
    assert used and used <= declared, used - declared
    methods = set(re.findall(r"pub fn ([a-z_0-9]+)", lib_rs))
    called = set(re.findall(r"self\.env\.([a-z_0-9]+)\(", scalar_rs)) | set(re.findall(r"GpuVecEnv::([a-z_0-9]+)\(", scalar_rs))
    assert called and called <= methods, called - methods
    assert "pub mod scalar;" in lib_rs and "track_returns" in sys_rs
