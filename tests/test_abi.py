"""CPU checks of the drop-in boundary: libdogm_b200.so loads without a GPU and exports every symbol that
include/dogm_b200.h declares; layouts match the reference's PODs; no compute is called here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from _loader import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dogm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dogm_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol(dogm_b200):
    lib = dogm_b200.load_library()  # raises when the .so is missing: there is no CPU fallback
    declared = declared_symbols()
    assert len(declared) >= 60
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/dogm_b200.h but not exported"
    # the binding covers exactly the declared set
    assert sorted(dogm_b200.EXPORTED_SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", dogm_b200.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (dogm_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_pod_layouts_match_reference(dogm_b200):
    # GridCell 64 B, MeasurementCell 16 B, Particle 28 B, Params 11 x 4 B (dogm_types.h:13-49, dogm.h:26-60)
    assert dogm_b200.GRID_CELL_DTYPE.itemsize == 64
    assert dogm_b200.GRID_CELL_DTYPE.names[:2] == ("start_idx", "end_idx")
    assert dogm_b200.GRID_CELL_DTYPE.fields["covar_xy_vel"][1] == 60
    assert dogm_b200.MEAS_CELL_DTYPE.names == ("free_mass", "occ_mass", "likelihood", "p_A")
    assert ctypes.sizeof(dogm_b200.Params) == 44
    assert ctypes.sizeof(dogm_b200.LaserSensorParams) == 16
    p = dogm_b200.ParticlesSoA(5)
    assert p.block.size == 5 * 28
    p.state[3] = (1, 2, 3, 4)
    p.grid_cell_idx[2] = 7
    p.weight[4] = 0.5
    p.associated[1] = 1
    assert p.block[3 * 16 : 3 * 16 + 16].view("<f4").tolist() == [1, 2, 3, 4]
    assert p.block[5 * 16 + 2 * 4 : 5 * 16 + 3 * 4].view("<i4")[0] == 7
    assert p.block[5 * 20 + 4 * 4 : 5 * 20 + 5 * 4].view("<f4")[0] == 0.5
    assert p.block[5 * 24 + 1] == 1


def test_version_and_no_device_behaviour(dogm_b200):
    lib = dogm_b200.load_library()
    assert b"sm_100a" in lib.dogm_b200_version()
    if dogm_b200.device_count() == 0:
        # without a CUDA device construction fails loudly (cudaError), it does not fall back to anything
        params = dogm_b200.Params(10.0, 1.0, 2, 1, 0.5, 0.0, 0.0, 0.02, 10.0, 30.0, 0.01)
        with pytest.raises(dogm_b200.DogmError):
            dogm_b200.DOGM(params)


def test_invalid_arguments_are_rejected(dogm_b200):
    lib = dogm_b200.load_library()
    h = ctypes.c_void_p()
    assert lib.dogm_create(None, ctypes.byref(h)) == -1
    bad = dogm_b200.Params(10.0, 0.0, 2, 1, 0.5, 0.0, 0.0, 0.02, 10.0, 30.0, 0.01)
    assert lib.dogm_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert lib.dogm_update_grid(None, None, 0.0, 0.0, 0.0, 0.1, 1) == -1
    assert lib.dogm_get_grid_size(None) == 0


def test_product_does_not_touch_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "dynamic-occupancy-grid-map_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("there is no oracle", ""), os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libdogm_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "dogm_ref" not in out
