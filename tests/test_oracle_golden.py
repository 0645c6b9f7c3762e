"""Pins the CPU oracle against outputs of the reference's own CUDA implementation (tests/golden/ref_*.npz, produced on a
B200 by oracle/gen_golden.py from oracle/_ref).  Runs on the CPU; needs neither a GPU nor /root/reference."""
import glob
import os

import numpy as np
import pytest

from _loader import ROOT
from _refcheck import OracleAdapter, check_cycle, check_first_cycle_init

FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_*.npz")))


def params_from(mod, arr):
    return mod.Params(float(arr[0]), float(arr[1]), int(arr[2]), int(arr[3]), *[float(v) for v in arr[4:]])


def test_golden_fixtures_exist():
    assert FIXTURES, "tests/golden/ref_*.npz missing: run oracle/gen_golden.py on a GPU box"


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_replays_reference_run(orc, path):
    z = np.load(path)
    params = params_from(orc, z["params"])
    impl = OracleAdapter(orc, params)
    gs = impl.o.grid_size
    dt = float(z["dt"][0])
    for c in range(int(z["cycles"][0])):
        snap = {k[len(f"c{c}_"):]: z[k] for k in z.files if k.startswith(f"c{c}_")}
        meas, pose = snap.pop("meas"), snap.pop("pose")
        for k in ("G0", "G2", "G3", "G4", "G5", "G6"):
            snap[k] = snap[k].view(orc.GRID_CELL_DTYPE)
        if c == 0:
            check_first_cycle_init(OracleAdapter(orc, params), snap, meas, impl.N, gs)
        stats = check_cycle(impl, snap, meas, float(pose[0]), float(pose[1]), float(pose[2]), dt, first_cycle=(c == 0))
        print(os.path.basename(path), "cycle", c, stats)
