"""Per-kernel CUDA-event times of the resident loop (`value`) and of the end-to-end loop (`e2e`) side by side (tools only).
    python tools/e2e_breakdown.py [config] [cycles]"""
import ctypes as C_
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
lib = gpu.load_library()
name = sys.argv[1] if len(sys.argv) > 1 else "nuss"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cfg = bench.CONFIGS[name]
beams = bench.make_beams(cfg, 64, seed=1234)
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
d = gpu.DOGM(params)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
ptr = gen.generate_grid(beams[0])
step = 0


def resident(sync):
    global step
    x, y = bench.pose_at(step)
    d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True, sync=sync)
    step += 1


beam_pinned = gpu.pinned_empty((cfg["beams"],), np.float32)
cap = 1 << 16
out = gpu.pinned_empty((cap,), gpu.DYNAMIC_CELL_DTYPE)
cnt = C_.c_int(0)


def e2e():
    global step
    beam_pinned[:] = beams[step % len(beams)]
    gen.generate_grid_into(d, beam_pinned)
    x, y = bench.pose_at(step)
    d.update_grid(None, float(x), float(y), 0.0, bench.DT, sync=False)
    if os.environ.get("E2E_NOFILTER"):
        d.synchronize()
    else:
        lib.dogm_extract_dynamic_cells(d._h, 0.7, 4.0, C_.c_void_p(out.ctypes.data), cap, C_.byref(cnt))
    step += 1


def timed(fn, label):
    for _ in range(12):
        fn()
    d.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        fn()
    d.synchronize()
    wall = (time.perf_counter() - t0) / K * 1e6
    d.kernel_timing_enable(True)
    for _ in range(K):
        fn()
    d.synchronize()
    kt = d.kernel_timing_read()
    d.kernel_timing_enable(False)
    print(f"{label}: {wall:.1f} us per cycle (wall, free running)")
    return {k: v["total_ms"] / K * 1e3 for k, v in kt.items()}


a = timed(lambda: resident(False), "resident")
if not os.environ.get("E2E_NOFILTER"):
    d.set_dynamic_cell_filter(0.7, 4.0, cap)
b = timed(e2e, "e2e")
print(f"{'kernel':24s} {'resident':>9s} {'e2e':>9s}   (us per cycle, every launch bracketed by events)")
for k in sorted(set(a) | set(b), key=lambda k: -(b.get(k, 0.0))):
    print(f"{k:24s} {a.get(k, 0.0):9.1f} {b.get(k, 0.0):9.1f}")
print(f"{'sum':24s} {sum(a.values()):9.1f} {sum(b.values()):9.1f};  dynamic cells in the last list: {cnt.value}")
