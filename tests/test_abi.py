"""The C-ABI library loads and exports every symbol include/mctomo_b200.h declares; without a GPU it
fails loudly instead of falling back; the product never references the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as orc
from mctomo_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "mctomo_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(mct_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported():
    names = _declared_symbols()
    assert len(names) >= 20
    L = C.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_integration_index_lists_every_entry_point():
    """INTEGRATION.md section 8 (entry point -> reference lines it replaces) names exactly the header's symbols."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read().split("## 8. Entry-point index")[1]
    listed = set(re.findall(r"`(mct_[a-z0-9_]+)`", doc))
    declared = set(_declared_symbols())
    assert listed == declared, (sorted(declared - listed), sorted(listed - declared))


def test_binding_loads():
    assert capi.lib() is not None and capi._bind_batch() is not None


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    L = capi.lib()
    rc = L.mct_init(0)
    assert rc == capi.MCT_E_CUDA and b"no CPU fallback" in L.mct_last_error()
    with pytest.raises(capi.MctError) as e:
        capi.vs2vp_rho(np.ones(8))
    assert e.value.code == capi.MCT_E_NOINIT


def test_box_window_matches_reference_formula():
    """mct_box_window is pure host logic (src/mcmc_loc2.f90:2034-2045) and needs no device."""
    rng = np.random.default_rng(0)
    grid = synth.make_grid(41, 37, 23)
    for _ in range(200):
        a = rng.uniform([-7, -7, -2], [7, 7, 14], (2, 3))
        box = np.concatenate([a.min(0), a.max(0)])
        assert np.array_equal(capi.box_window(grid, box), orc.box_window(grid, box))
    w = capi.box_window(grid, grid.full_box())
    assert list(w) == [1, 41, 1, 37, 1, 23]


def test_product_does_not_reference_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "mctomo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"liboracle|import\s+oracle|from\s+oracle|include\s+\"[^\"]*oracle/", src):
                    bad.append(f)
    assert not bad, f"product sources reference the oracle: {bad}"
    import subprocess
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_slab_bounds_is_host_logic_and_covers_the_grid():
    """mct_slab_bounds needs no device; slabs are contiguous, equal width, cover 1..nx once; same rule as shard.py."""
    from mctomo_b200 import shard
    for nx in (1, 5, 7, 8, 64, 1000, 1024):
        for world in (1, 2, 3, 8):
            cols = []
            for r in range(world):
                lo, hi, per = capi.slab_bounds(nx, world, r)
                assert (lo, hi, per) == shard.slab_bounds(nx, world, r)
                cols += list(range(lo, hi + 1))
            assert cols == list(range(1, nx + 1))


def test_comm_without_device_fails_loudly_and_library_has_no_nccl_dependency():
    import subprocess
    import torch
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in out                       # bound with dlopen at mct_comm_init, not at link time
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    L = capi.lib()
    assert L.mct_comm_init(b"\0" * 128, 0, 2) == capi.MCT_E_NOINIT
    assert capi.comm_info()["active"] is False
