"""Band mode over the GPUs of one box (tools only): the large grids of BASELINE.json split into bands of rows, one band per
GPU, particles migrating between neighbours by peer copies.  Prints ms per cycle for 1, 2, 4, ... GPUs (as many as visible).
    python tools/band_bench.py [config] [cycles]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from _loader import load_dogm_b200  # noqa: E402

gpu = load_dogm_b200()
name = sys.argv[1] if len(sys.argv) > 1 else "highway"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = bench.CONFIGS[name]
ndev = gpu.device_count()
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
beams = bench.make_beams(cfg, int(os.environ.get("BAND_SCANS", "4")), seed=1234)
gpu.set_device(0)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
grids = [gen.generate_grid_host(b) for b in beams]  # full grids on the host; every band gets its rows
gen.close()
# bands of about equal load.  Where the particles sit is a property of the scene, so it is measured: the single-band run
# leaves a row histogram of its particles, which sets the band edges of the multi-band runs (every row also costs its cells)
G = int(np.sqrt(grids[0].size))
row_load = None
row_hist = None
want = [int(v) for v in os.environ.get("BAND_COUNTS", "1,2,4,8").split(",")]  # the 1-band run also measures the row loads
# BAND_EDGES: "phase" = edges that minimise the sum over the phases of the slowest band (balanced_rows_by_phase),
#             "load"  = one load figure per row (balanced_rows); both can be listed
methods = os.environ.get("BAND_EDGES", "phase").split(",")
runs = [(r, m) for r in want if r <= ndev for m in (methods if r > 1 else methods[:1])]
for R, method in runs:
    rows = None
    if row_load is not None and R > 1:
        rows = gpu.balanced_rows_by_phase(row_hist, R) if method == "phase" else gpu.balanced_rows(row_load, R)
    paced = {"host": False, "device": True}.get(os.environ.get("BAND_PACED", ""), None)
    bd = gpu.BandedDOGM(params, R, devices=list(range(R)), seed=123456, rows=rows, slack=2.5, device_paced=paced)
    meas = []  # per band: its rows of every scan, resident on the band's GPU
    for r in range(R):
        gpu.set_device(r)
        lo, hi = bd.row0[r] * bd.G, (bd.row0[r] + bd.rows[r]) * bd.G
        ptrs = []
        for g in grids:
            part = np.ascontiguousarray(g[lo:hi])
            p = gpu.device_alloc(part.nbytes)
            gpu.memcpy_h2d(p, part)
            ptrs.append(p)
        meas.append(ptrs)
    step = 0

    def cycle():
        global step
        x, y = bench.pose_at(step)
        counts = bd.update_grid([meas[r][step % len(grids)] for r in range(R)], float(x), float(y), 0.0, bench.DT)
        step += 1
        return counts

    for _ in range(6):
        counts = cycle()
    t0 = time.perf_counter()
    for _ in range(K):
        counts = cycle()
    t = (time.perf_counter() - t0) / K
    lo, hi = bd.last_migration
    print(f"{name}: {R} band(s) on {R} GPU(s), {'device' if bd.device_paced else 'host'}-paced, edges by {method if R > 1 else '-'}, rows {bd.rows}: {1e3 * t:8.3f} ms/cycle, {1.0 / t:8.1f} cycles/s; particles per band {counts}, "
          f"migrated last cycle {sum(lo) + sum(hi)}; phases [predict, exchange, update, birth+cdf, resample] ms "
          f"{[round(v, 2) for v in bd.last_phase_ms]}; peer access {bd.peer_access}")
    if getattr(bd, "last_band_ms", None):
        print("    per band [predict, exchange, update, birth+cdf, resample] ms:", [[round(v, 2) for v in row] for row in bd.last_band_ms])
    if row_load is None:
        hist = np.zeros(G, np.float64)
        for r in range(R):
            state = bd.get_particles(r)[0]
            hist += np.bincount(np.clip(state[:, 1].astype(np.int64), 0, G - 1), minlength=G)
            del state
        row_hist = hist
        row_load = hist / hist.sum() * (240.0 * cfg["n"] + 37.0 * cfg["b"]) + 100.0 * G  # measured: a particle costs about 2.4 cells
    for r in range(R):
        gpu.set_device(r)
        for p in meas[r]:
            gpu.device_free(p)
    bd.close()
