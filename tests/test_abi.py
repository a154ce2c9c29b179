"""CPU-only checks of the drop-in boundary: libclumpy_cuda.so loads, exports exactly what
include/clumpy_cuda.h declares, and its host-side helpers (no device needed) agree with the oracle.
No kernels are launched here.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import clumpy_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "clumpy_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clumpy_cuda_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 30
    assert sorted(clumpy_b200.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    L = clumpy_b200.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    raw = C.CDLL(clumpy_b200.LIB_PATH)
    for name in declared_symbols():
        getattr(raw, name)  # raises AttributeError if the export is missing


def test_no_torch_or_oracle_in_the_product():
    """The boundary is plain C; the product never reaches into oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "clumpy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), (dirpath, f)
    hdr = open(os.path.join(ROOT, "include", "clumpy_cuda.h")).read()
    assert "torch" not in hdr and "at::" not in hdr


def test_age_offsets_match_oracle_and_goldens(oracle_mod, golden):
    for nframes, first in golden["meta"]["age_offsets_first10"].items():
        assert clumpy_b200.age_offsets(int(nframes), 10).tolist() == first
    for nframes in (1, 2, 3, 240, 1000, 65537):
        a = clumpy_b200.age_offsets(nframes, 5000)
        assert np.array_equal(a, oracle_mod.age_offsets(nframes, 5000)), nframes
    # the library draws a whole generator block at a time and only falls back to draw-by-draw when a block
    # contains a rejection candidate: exercise block boundaries, ragged tails and rejection-heavy ranges
    for nframes, n in ((1000, 624), (1000, 625), (1000, 1 << 20), (1000003, 300001), (1500000000, 300000),
                       (2**31 - 1, 200000), (2**30 + 7, 200000), (2**31, 100000)):
        assert np.array_equal(clumpy_b200.age_offsets(nframes, n), oracle_mod.age_offsets(nframes, n)), (nframes, n)


def test_simplex_perm_matches_oracle(oracle_mod, golden):
    assert clumpy_b200.simplex_perm(0)[:8].tolist() == golden["meta"]["perm_seed0_first8"]
    for seed in (0, 1, -1, 26, 2**31 - 1, -(2**31)):
        assert np.array_equal(clumpy_b200.simplex_perm(seed), oracle_mod.simplex_perm(seed)), seed


def test_invalid_arguments_fail_without_a_device():
    L = clumpy_b200.lib()
    assert L.clumpy_cuda_age_offsets(0, 4, np.zeros(4, np.uint32).ctypes.data_as(C.POINTER(C.c_uint32))) == clumpy_b200.ERR_INVALID
    assert b"nframes" in L.clumpy_cuda_last_error()
    assert L.clumpy_cuda_sync(None) == clumpy_b200.ERR_INVALID


def test_no_silent_cpu_fallback():
    """Without a CUDA device the context cannot be created; nothing computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(clumpy_b200.ClumpyCudaError):
        clumpy_b200.Context(0)


def test_both_build_files_list_every_source():
    """The Makefile (what build() runs) and CMakeLists.txt (the sm_100a CMake target) compile the same sources."""
    csrc = os.path.join(ROOT, "clumpy_b200", "csrc")
    sources = sorted(f for f in os.listdir(csrc) if f.endswith((".cu", ".cc")))
    make = open(os.path.join(ROOT, "Makefile")).read()
    cmake = open(os.path.join(ROOT, "CMakeLists.txt")).read()
    for f in sources:
        stem = f.rsplit(".", 1)[0]
        assert f in cmake, f
        assert f in make or f"{stem}.o" in make, f
    assert "CMAKE_CUDA_ARCHITECTURES 100a" in cmake and "compute_100a,code=sm_100a" in make
