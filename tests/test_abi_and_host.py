"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every symbol the header
declares; creating a context without a GPU fails loudly (no CPU fallback); host-side helpers."""
import ctypes
import os
import re

import numpy as np
import pytest

import xpic_b200 as X
from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    X.build_library()
    return X.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "xpic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xb_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = header_symbols()
    assert len(names) >= 30
    raw = ctypes.CDLL(X.library_path())
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing
    # and the Python binding covers the header one to one
    from xpic_b200.binding import SYMBOLS

    assert sorted(SYMBOLS) == names


def test_stencil_table(lib):
    t = X.coef_table()
    assert t.shape == (369, 5)
    for c1 in range(3):
        for c2 in range(3):
            assert np.sum((t[:, 0] == c1) & (t[:, 1] == c2)) == (27 if c1 == c2 else 48)
    assert len({tuple(r) for r in t}) == 369
    # the 13 non-zeros per row of M = 2I + dt^2/2 curl curl are a subset of the slots
    slots = {tuple(r) for r in t}
    for c in range(3):
        assert (c, c, 0, 0, 0) in slots
        for a in range(3):
            if a == c:
                continue
            e = [0, 0, 0]
            e[a] = 1
            assert (c, c, *e) in slots and (c, c, *[-v for v in e]) in slots
            ec = [0, 0, 0]
            ec[c] = 1
            for off in ([0, 0, 0], ec, [-v for v in e], [ec[i] - e[i] for i in range(3)]):
                assert (c, a, *off) in slots


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(X.XpicB200Error, match="no CUDA device"):
        X.Simulation((8, 8, 8))


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xpic_b200")):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), (dirpath, f)  # oracle/ is test infrastructure only


def test_slab_ranges_tile_the_box():
    for nz in (8, 13, 128, 257):
        for nranks in (1, 2, 3, 4, 8):
            if nz < nranks:
                continue
            z = 0
            for r in range(nranks):
                z0, nzl = X.slab_range(nz, r, nranks)
                assert z0 == z and nzl >= 1
                for k in range(z0, z0 + nzl):
                    assert X.owner_rank(k, nz, nranks) == r
                z += nzl
            assert z == nz


def test_accumulator_tile_map_matches_particle_footprint():
    """The compile-time map of the moment deposition's accumulator tiles (csrc/deposit.cuh: tile element -> node and
    stencil slot, used by k_gather_tiles) against the footprint of a particle as the reference deposits it
    (src/impls/ecsim/particles.cpp:119-171), for every component pair, octant, tile row and column.  Host code only."""
    import subprocess

    exe = os.path.join(ROOT, "xpic_b200", "_build", "tile_map_check")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "xpic_b200", "csrc"), "../_build/tile_map_check"], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
