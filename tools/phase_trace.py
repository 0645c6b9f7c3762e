"""Phase timeline inside k_cdf_chain and k_resample (tools only; needs a -DDOGM_PHASE_TRACE build:
    tools/build_variant.sh phase -DDOGM_PHASE_TRACE;  DOGM_B200_LIB=build_ab/lib_phase.so python tools/phase_trace.py [config]
Thread 0 of every CTA stamps %globaltimer at its phase boundaries; printed are, per phase boundary, the time since the
kernel's first stamp: min / median / max over the CTAs."""
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
beams = bench.make_beams(cfg, 8, seed=1234)
params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
d = gpu.DOGM(params)
gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
ptr = gen.generate_grid(beams[0])
for step in range(14):
    x, y = bench.pose_at(step)
    d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True, sync=False)
d.synchronize()
n_chain = (cfg["n"] + cfg["b"] + 4095) // 4096
n_res = (cfg["n"] + 1023) // 1024
LABELS = {
    "phase_chain": ["ticket", "entries loaded", "block total", "prefix known", "cdf written", "total known", "claims done"],
    "phase_resample": ["start", "window staged", "claim checked", "ancestors", "slots", "records", "stored"],
}
LABELS["phase_birth"] = ["start", "scale known", "owner found", "weights + noise", "stored", "-", "-"]
for key, n in (("phase_chain", n_chain), ("phase_resample", n_res), ("phase_birth", (cfg["b"] + 255) // 256)):
    buf = np.zeros((4096, 8), np.uint64)
    d.debug_read(key, buf)
    t = buf[: min(n, 4096), :7].astype(np.int64)
    last = 4 if key == "phase_birth" else 6
    t[:, last + 1:] = t[:, last:last + 1]
    t0 = t[:, 0].min()
    rel = (t - t0) * 1e-3
    print(f"{key}: {n} CTAs, kernel span {rel.max():.1f} us (first stamp to last)")
    for k, lab in enumerate(LABELS[key]):
        c = rel[:, k]
        print(f"   {lab:16s} min {c.min():7.1f}  p10 {np.percentile(c, 10):7.1f}  median {np.median(c):7.1f}  p90 {np.percentile(c, 90):7.1f}  max {c.max():7.1f}")
    dur = rel[:, last] - rel[:, 0]
    print(f"   per-CTA duration: median {np.median(dur):.1f} us, max {dur.max():.1f} us; CTA start times: median {np.median(rel[:,0]):.1f}, p90 {np.percentile(rel[:,0],90):.1f}, max {rel[:,0].max():.1f}")
    steps = np.diff(rel, axis=1)
    print("   median step durations:", " ".join(f"{LABELS[key][k+1]}={np.median(steps[:,k]):.2f}" for k in range(6)))
