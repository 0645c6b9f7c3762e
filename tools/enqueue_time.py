"""Host-side cost of enqueueing DOGM cycles (tools only): times 20 asynchronous cycles without the final synchronise."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from _loader import load_dogm_b200  # noqa: E402

gpu = load_dogm_b200()
name = sys.argv[1] if len(sys.argv) > 1 else "nuss"
cfg = bench.CONFIGS[name]
beams = bench.make_beams(cfg, 8, seed=1234)
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
d = gpu.DOGM(params)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
ptr = gen.generate_grid(beams[0])
step = 0
for _ in range(10):
    x, y = bench.pose_at(step)
    d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True)
    step += 1
for rep in range(3):
    d.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        x, y = bench.pose_at(step)
        d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True, sync=False)
        step += 1
    t1 = time.perf_counter()
    d.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: enqueue {1e6 * (t1 - t0) / 20:.1f} us/cycle, total {1e6 * (t2 - t0) / 20:.1f} us/cycle")
