import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from _loader import load_dogm_b200
os.environ["DOGM_B200_SKIP"] = "0x2000"   # no k_resample from cycle 12 on: the claims stay in res_start
gpu = load_dogm_b200()
cfg = bench.CONFIGS["nuss"]
beams = bench.make_beams(cfg, 8, seed=1234)
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
d = gpu.DOGM(params)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
ptr = gen.generate_grid(beams[0])
for step in range(13):
    x, y = bench.pose_at(step)
    d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True)
N = cfg["n"]
rs = d.debug_read("res_start", np.empty((N + 255) // 256, np.int32))
tot = d.debug_read("weight_total", np.empty(1, np.float64))[0]
cdf = d.get_joint_weight_accum()
print("total", tot, "cdf[-1]", cdf[-1], "n", cdf.size)
print("claimed", int((rs < cdf.size).sum()), "of", rs.size, "first", rs[:8])
