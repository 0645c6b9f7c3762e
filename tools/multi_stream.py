"""Several independent DOGM handles (sensor streams) on one GPU, each on its own CUDA stream (BASELINE.json configs[3]:
64 lidar streams over 8 GPUs = 8 per GPU).  Prints scenes*cycles/s for 1, 2, 4, 8 concurrent handles (tools only).
    python tools/multi_stream.py [config] [cycles]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from _loader import load_dogm_b200

gpu = load_dogm_b200()
name = sys.argv[1] if len(sys.argv) > 1 else "nuss"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cfg = bench.CONFIGS[name]
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
C = None
for S in (1, 2, 4, 8):
    handles, rings = [], []
    for s in range(S):
        d = gpu.DOGM(params)
        d.set_options(seed=1000 + s, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
        beams = bench.make_beams(cfg, 4, seed=77 + s)
        ptrs = []
        for bm in beams:
            src = gen.generate_grid(bm)
            host = np.empty(d.grid_cell_count, dtype=gpu.MEAS_CELL_DTYPE)
            gpu.memcpy_d2h(host, src)
            p = gpu.device_alloc(host.nbytes)
            gpu.memcpy_h2d(p, host)
            ptrs.append(p)
        handles.append(d)
        rings.append(ptrs)
    step = 0

    def round_(sync):
        global step
        x, y = bench.pose_at(step)
        for d, ring in zip(handles, rings):
            d.update_grid(ring[step % len(ring)], float(x), float(y), 0.0, bench.DT, device=True, sync=False)
        step += 1
        if sync:
            for d in handles:
                d.synchronize()

    for _ in range(10):
        round_(True)
    t0 = time.perf_counter()
    for _ in range(K):
        round_(False)
    for d in handles:
        d.synchronize()
    t = time.perf_counter() - t0
    print(f"{name}: {S} concurrent handles: {S * K / t:9.1f} scenes*cycles/s  ({1e3 * t / K:.3f} ms per round of {S})")
    for d, ring in zip(handles, rings):
        for p in ring:
            gpu.device_free(p)
        d.close()
gen.close()
