"""The reference's demo loop (demo/main.cpp:22-103) end to end: simulated lidar scene -> measurement grid -> DOGM cycle ->
dynamic cells -> DBSCAN -> MAE / RMSE against the simulated vehicles, for this implementation and, when oracle/_ref is
present, for the reference's CUDA code on the same scans (tools only).
    python tools/demo_quality.py [particles] [birth] [alt]"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _loader import PKG_DIR, load_dogm_b200, load_oracle, load_ref  # noqa: E402

spec = importlib.util.spec_from_file_location("dogm_b200_tools", os.path.join(PKG_DIR, "tools.py"))
tools_mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tools_mod)

DEMO = dict(num_points=100, fov=120.0, grid_size=50.0, ego_velocity=(0.0, 4.0), steps=14, dt=0.1)
PARAMS = (0.99, 0.1, 1.0, 0.02, 30.0, 30.0, 0.01)  # demo/main.cpp:28-34
RES = 0.2


def run_mine(gpu, t, vehicles, n, b, seed=123456):
    meas, _, ego = t.simulate(vehicles=vehicles, **DEMO)
    d = gpu.DOGM(gpu.Params(50.0, RES, n, b, *PARAMS))
    d.set_options(seed=seed, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(50.0, RES, 120.0, 0.5), 50.0, RES)
    ev = t.evaluator(vehicles=vehicles, resolution=RES, **DEMO)
    d.set_dynamic_cell_filter(0.7, 4.0)
    counts = []
    for s in range(DEMO["steps"]):
        ptr = gen.generate_grid(meas[s])
        d.update_grid(ptr, float(ego[s, 0]), float(ego[s, 1]), 0.0, DEMO["dt"], device=True)
        cells, count = d.extract_dynamic_cells(0.7, 4.0)
        counts.append(count)
        ev.step(s, cells, d.grid_size)
    out = ev.summary()
    out["dynamic_cells"] = counts
    ev.close()
    gen.close()
    d.close()
    return out


def run_reference(gpu, t, vehicles, n, b):
    orc, ref = load_oracle(), load_ref()
    meas, _, ego = t.simulate(vehicles=vehicles, **DEMO)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(50.0, RES, 120.0, 0.5), 50.0, RES)
    r = ref.RefDOGM(orc.Params(50.0, RES, n, b, *PARAMS), orc.GRID_CELL_DTYPE, orc.MEAS_CELL_DTYPE)
    ev = t.evaluator(vehicles=vehicles, resolution=RES, **DEMO)
    counts = []
    for s in range(DEMO["steps"]):
        grid = gen.generate_grid_host(meas[s])  # the reference's own GL path cannot run here: same measurement grid for both
        r.update_grid(grid.view(orc.MEAS_CELL_DTYPE), float(ego[s, 0]), float(ego[s, 1]), 0.0, DEMO["dt"], device=False)
        rec, count = orc.extract_dynamic_cells(r.get_grid_cells(), 0.7, 4.0)  # computeCellsWithVelocity restated (CPU)
        cells = np.zeros(count, tools_mod.DYNAMIC_CELL_DTYPE)
        cells["cell_idx"] = rec[:count, 0].copy().view(np.int32)
        cells["mean_x_vel"], cells["mean_y_vel"] = rec[:count, 2], rec[:count, 3]
        counts.append(count)
        ev.step(s, cells, 250)
    out = ev.summary()
    out["dynamic_cells"] = counts
    ev.close()
    gen.close()
    r.close()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
    vehicles = tools_mod.DEMO_VEHICLES_ALT if len(sys.argv) > 3 else tools_mod.DEMO_VEHICLES
    gpu, t = load_dogm_b200(), tools_mod.Tools()
    print("this implementation:", run_mine(gpu, t, vehicles, n, b))
    if load_ref().available():
        print("reference CUDA code:", run_reference(gpu, t, vehicles, n, b))
