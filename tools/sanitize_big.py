"""A few cycles at a larger size for compute-sanitizer (tools only): python tools/sanitize_big.py [config] [cycles]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from _loader import load_dogm_b200
gpu = load_dogm_b200()
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "demo"]
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 3
beams = bench.make_beams(cfg, cycles, seed=5)
d = gpu.DOGM(gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS))
gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE), cfg["size"], cfg["resolution"])
d.set_dynamic_cell_filter(0.7, 4.0)
for s in range(cycles):
    ptr = gen.generate_grid(beams[s])
    x, y = bench.pose_at(s)
    d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True)
    d.extract_dynamic_cells(0.7, 4.0)
print("done", float(d.get_grid_cells()["occ_mass"].sum()))
