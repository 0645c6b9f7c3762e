"""Runs a few steady-state DOGM cycles of the headline configuration for ncu (tools only; not part of the product).

    ncu --set full --clock-control none --import-source on --launch-skip <S> -c <N> -o gpurun_out/prof python tools/profile_cycle.py
    python tools/profile_cycle.py --print-skip     # number of kernel launches before the profiled cycles
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from _loader import load_dogm_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="nuss")
    ap.add_argument("--warm", type=int, default=12)
    ap.add_argument("--cycles", type=int, default=2)
    ap.add_argument("--print-skip", action="store_true")
    args = ap.parse_args()
    gpu = load_dogm_b200()
    cfg = bench.CONFIGS[args.config]
    beams = bench.make_beams(cfg, args.warm + args.cycles + 2, seed=1234)
    params = gpu.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
    laser = gpu.LaserSensorParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
    d = gpu.DOGM(params)
    gen = gpu.LaserMeasurementGrid(laser, cfg["size"], cfg["resolution"])
    step = 0
    for _ in range(args.warm):
        ptr = gen.generate_grid(beams[step])
        x, y = bench.pose_at(step)
        d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True)
        step += 1
    # generate_grid launches k_meas_polar + k_meas_apply per call and k_meas_geom once (not counted by the DOGM handle)
    skip = d.launch_count() + 2 * args.warm + 1
    if args.print_skip:
        print(skip)
        return
    for _ in range(args.cycles):
        ptr = gen.generate_grid(beams[step])
        x, y = bench.pose_at(step)
        d.update_grid(ptr, float(x), float(y), 0.0, bench.DT, device=True)
        step += 1
    print("profiled cycles done; launches before them:", skip, "per cycle:", (d.launch_count() - (skip - 2 * args.warm - 1)) // args.cycles)


if __name__ == "__main__":
    main()
