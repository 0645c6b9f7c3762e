"""Developer tool: bucket sizes of the bucket sort on the headline scene (after a few cycles)."""
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
cfg = bench.CONFIGS[name]
beams = bench.make_beams(cfg, 16, seed=1234)
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
os.environ["DOGM_B200_SORT"] = "bucket"
d = gpu.DOGM(params)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
for step in range(14):
    ptr = gen.generate_grid(beams[step])
    x, y = bench.pose_at(step)
    d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True)
C = d.grid_cell_count
shift = 11 if cfg["n"] <= 1024 * 2048 else 12
tiles = (cfg["n"] + (1 << shift) - 1) >> shift  # splitter samples
bins = ((tiles + (C + 2047) // 2048 + 1) + 31) // 32 * 32
start = d.debug_read("bucket_start", np.zeros(bins, np.uint32)).astype(np.int64)
org = d.debug_read("bucket_org", np.zeros(bins, np.int32))
smp = d.debug_read("bucket_samples", np.zeros(tiles, np.int32))
size = np.diff(np.concatenate([start, [cfg["n"]]]))
print("bins", bins, "non-empty", int((size > 0).sum()), "max", int(size.max()), "p99", np.percentile(size[size > 0], 99), "median", np.median(size[size > 0]))
print("sizes > 4096:", int((size > 4096).sum()), "sum of those", int(size[size > 4096].sum()), "largest 10:", np.sort(size)[-10:])
print("histogram of sizes (k):", np.histogram(size[size > 0], bins=[1, 512, 1024, 2048, 3072, 4096, 6144, 8192, 16384, 1 << 30])[0])
print("samples monotone violations:", int((np.diff(smp) < 0).sum()), "equal neighbours:", int((np.diff(smp) == 0).sum()))
p = d.get_particles()
idx = p.grid_cell_idx
print("population sortedness: descents in cell idx:", int((np.diff(idx) < 0).sum()))
cnt = np.bincount(idx, minlength=C)
print("particles per occupied cell: max", cnt.max(), "p99", np.percentile(cnt[cnt > 0], 99), "cells>4096:", int((cnt > 4096).sum()), "occupied", int((cnt > 0).sum()))
