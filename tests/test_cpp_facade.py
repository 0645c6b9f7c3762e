"""The drop-in C++ class (include/dogm/dogm.h over the C ABI): compiles everywhere, runs the restated dogm_spec.cpp
cases on the GPU box."""
import os
import subprocess

import pytest

from _loader import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "dogm_spec_b200.cpp")
PKG = os.path.join(ROOT, "dynamic-occupancy-grid-map_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "dogm_spec_b200")


def build():
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", PKG, "-ldogm_b200", f"-Wl,-rpath,{PKG}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return EXE


def test_facade_compiles_with_a_plain_cxx14_compiler():
    build()  # the header-only class needs neither nvcc nor GLM nor the CUDA headers
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_restated_dogm_spec_cases_pass_on_the_gpu():
    exe = build()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all passed" in res.stdout
