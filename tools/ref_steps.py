"""Per-step wall-clock times of the reference's updateGrid (oracle/_ref) at the headline configuration (tools only)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from _loader import load_oracle, load_ref
orc, ref = load_oracle(), load_ref()
cfg = bench.CONFIGS["nuss"]
beams = bench.make_beams(cfg, 8, seed=1234)
laser = orc.LaserParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
p = orc.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
r = ref.RefDOGM(p, orc.GRID_CELL_DTYPE, orc.MEAS_CELL_DTYPE)
ring = [ref.device_upload(orc.meas_generate(laser, cfg["size"], cfg["resolution"], bm)) for bm in beams]
ts = []
for step in range(40):
    x, y = bench.pose_at(step)
    t0 = time.perf_counter()
    r.update_grid(ring[step % len(ring)], float(x), float(y), 0.0, bench.DT, device=True)
    ts.append((time.perf_counter() - t0) * 1e3)
print("per-step ms:", " ".join(f"{t:.1f}" for t in ts))
