"""Generates tests/golden/ref_*.npz: inputs, extracted noise and per-stage outputs of the REFERENCE's own CUDA code
(oracle/_ref, built from /root/reference by oracle/build_ref.sh) on small configurations.

Run on a GPU box:   python oracle/gen_golden.py --out gpurun_out/golden
and copy the files into tests/golden/.  TEST INFRASTRUCTURE ONLY.  The fixtures pin the CPU oracle
(tests/test_oracle_golden.py) without needing a GPU or /root/reference at test time."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _loader import load_oracle, load_ref  # noqa: E402
from conftest import make_params, synthetic_meas  # noqa: E402

CASES = {
    # name: (size, resolution, N, B, cycles, ego velocity (x, y) per cycle in metres, dt, seed)
    "small": (16.0, 0.5, 4096, 512, 3, (0.0, 0.6), 0.1, 101),
    "ragged": (9.0, 0.3, 1001, 77, 3, (-0.4, 0.35), 0.1, 102),
    "spec": (10.0, 1.0, 2, 1, 3, (3.0, 0.0), 0.1, 103),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    orc, ref = load_oracle(), load_ref()
    for name, (size, res, n, b, cycles, ego, dt, seed) in CASES.items():
        rng = np.random.default_rng(seed)
        params = make_params(orc, size, res, n, b)
        r = ref.RefDOGM(params, orc.GRID_CELL_DTYPE, orc.MEAS_CELL_DTYPE)
        out = {
            "params": np.array([size, res, n, b, params.persistence_prob, params.stddev_process_noise_position,
                                params.stddev_process_noise_velocity, params.birth_prob, params.stddev_velocity,
                                params.init_max_velocity, params.freespace_discount], np.float64),
            "cycles": np.array([cycles]), "dt": np.array([dt], np.float32),
            "rng_threads": np.array([r._lib.ref_rng_thread_count(r._h)]),
        }
        for c in range(cycles):
            meas = synthetic_meas(orc.MEAS_CELL_DTYPE, r.grid_size, rng)
            x, y = np.float32(ego[0] * c), np.float32(ego[1] * c)
            snap = r.run_cycle_verbose(meas, float(x), float(y), 0.0, dt, first_cycle=(c == 0))
            out[f"c{c}_meas"] = meas
            out[f"c{c}_pose"] = np.array([x, y, 0.0], np.float32)
            for k, v in snap.items():
                out[f"c{c}_{k}"] = v
        r.close()
        path = os.path.join(args.out, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
