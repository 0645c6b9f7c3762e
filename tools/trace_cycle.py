"""Timeline of one free-running DOGM cycle from the library's own start stamps (tools only).
    python tools/trace_cycle.py [config] [repeats]
Prints, per kernel, the gap to the next kernel's start = its share of the cycle (median over the repeats)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from _loader import load_dogm_b200  # noqa: E402

gpu = load_dogm_b200()
name = sys.argv[1] if len(sys.argv) > 1 else "nuss"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
cfg = bench.CONFIGS[name]
beams = bench.make_beams(cfg, 8, seed=1234)
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
d = gpu.DOGM(params)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
ptrs = gen.generate_grid(beams[0])
step = 0


def cycle(sync):
    global step
    x, y = bench.pose_at(step)
    d.update_grid(ptrs, float(x), float(y), 0.0, bench.DT, device=True, sync=sync)
    step += 1


for _ in range(12):
    cycle(True)
d.trace_arm(True)
gaps = {}
order = []
for _ in range(reps):
    cycle(False)  # warm caches the way the steady state does
    d.synchronize()
    d.trace_read()
    cycle(False)
    d.extract_dynamic_cells(0.7, 4.0)  # one more (unchained) kernel: its stamp closes the last gap
    tr = d.trace_read()
    order = [k for k, _ in tr]
    for (k0, t0), (_, t1) in zip(tr[:-1], tr[1:]):
        gaps.setdefault(k0, []).append((t1 - t0) * 1e-3)
d.trace_arm(False)
total = 0.0
for k in order[:-1]:
    g = float(np.median(gaps[k]))
    total += g
    print(f"{k:22s} {g:8.1f} us")
print(f"{'sum':22s} {total:8.1f} us")
