"""CPU: the C-ABI library loads, exports every symbol include/vaxmc.h declares, and refuses
to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vaxmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(vax):
    nat = vax._native
    lib = nat.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vaxmc.h but not exported"
    assert sorted(nat.SYMBOLS) == declared, "ctypes table and header disagree"


def test_struct_layout_matches_header(vax):
    nat = vax._native
    # vx_config: 12 doubles + i64 + 2*i32 + i64 + u64 + 2 doubles + 2*i64 + double
    assert ctypes.sizeof(nat.vx_config) == 12 * 8 + 8 + 4 + 4 + 8 + 8 + 16 + 16 + 8
    assert ctypes.sizeof(nat.vx_result) == 12 * 8 + 8 + 8 + 16 + 4 + 4 + 8 + 8 + 8 + 8 + 8 + 3 * 8
    # ... and the built library agrees (the stamp _native.load() checks before handing the library out)
    lib = nat.load()
    assert lib.vx_abi_version() == nat.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "vaxmc.h")).read()
    assert int(re.search(r"#define VX_ABI_VERSION (\d+)", hdr).group(1)) == nat.ABI_VERSION
    # the direct/streamed crossover is decided in the library and mirrored by the Python driver
    assert int(re.search(r"#define VX_DIRECT_MAX_DEFAULT (\d+)", hdr).group(1)) == nat.DIRECT_MAX_DEFAULT
    assert lib.vx_sizeof(0) == ctypes.sizeof(nat.vx_config)
    assert lib.vx_sizeof(1) == ctypes.sizeof(nat.vx_result)
    assert lib.vx_sizeof(2) == ctypes.sizeof(nat.vx_nccl_id) == 128


def test_shard_range_matches_python_driver(vax):
    """vx_host_shard_range (what vx_run_bounds_comm shards by) == distributed.shard_range."""
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            for r in range(world):
                assert vax._native.shard_range(n, r, world) == vax.distributed.shard_range(n, r, world)


def test_library_is_sm100a_only(vax):
    import subprocess, shutil

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", vax._native.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback(vax):
    nat = vax._native
    assert nat.load().vx_device_count() == 0
    with pytest.raises(nat.VaxmcError) as e:
        nat.Context(0)
    assert e.value.code == nat.VX_ERR_NODEVICE
    with pytest.raises(nat.VaxmcError):
        vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95, 10, 50000,
                      dict(start_date="2020-12-15", end_date="2021-01-15"), seed=1)


def test_host_quantile_helpers_match_numpy(vax):
    lib = vax._native.load()
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 7, 100, 1001):
        x = np.sort(rng.integers(0, 1000, size=n)).astype(np.float64) * 0.37
        for q in (0.0, 0.025, 0.49525, 0.5, 0.50475, 0.975, 0.995, 1.0, 1 / 3):
            p, nx, g = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
            lib.vx_host_quantile_index(n, q, ctypes.byref(p), ctypes.byref(nx), ctypes.byref(g))
            got = lib.vx_host_lerp(x[p.value], x[nx.value], g.value)
            assert got == np.quantile(x, q), (n, q)


def test_host_count_to_value_matches_reference_formulas(vax):
    lib = vax._native.load()
    for N in (50000, 1000, 1000000, 333):
        for c in (0, 1, 17, 49999, 354315):
            assert lib.vx_host_count_to_value(0, c, N) == (c / N) * 100      # model.py:112,124
            assert lib.vx_host_count_to_value(1, c, N) == c * 1e6 / N        # model.py:125
            assert lib.vx_host_count_to_value(2, c, N) == c * 100 / N        # model.py:126
            assert lib.vx_host_count_to_value(3, c, N) == c * 100 / N        # model.py:127
    assert lib.vx_host_weeks(564) == 81 and lib.vx_host_weeks(1826) == 261
    assert lib.vx_host_rows(564) == 2 * 564 + 2 * 81
