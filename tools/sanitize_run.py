"""A short run of every code path for compute-sanitizer (tools only): small grid, ego shifts, stage API, both
resampling modes, dynamic-cell filter, measurement grid, read-outs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _loader import load_dogm_b200
from conftest import make_params
gpu = load_dogm_b200()
for (size, res, n, b) in [(20.0, 0.25, 20000, 2000), (13.0, 0.5, 5001, 333)]:
    p = make_params(gpu, size, res, n, b)
    d = gpu.DOGM(p)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(size, res, 120.0, 0.5), size, res)
    rng = np.random.default_rng(1)
    d.set_dynamic_cell_filter(0.6, 0.5, 4096)
    for mode in (gpu.RESAMPLE_SYSTEMATIC, gpu.RESAMPLE_STRATIFIED):
        d.set_options(seed=7, resample_mode=mode, noise_mode=gpu.NOISE_PHILOX)
        for c in range(4):
            z = np.where(rng.uniform(size=37) < 0.6, rng.uniform(2, size * 0.9, 37), np.inf).astype(np.float32)
            ptr = gen.generate_grid(z)
            d.update_grid(ptr, 0.3 * c, 0.45 * c, 0.0, 0.1, device=True)
            d.extract_dynamic_cells(0.6, 0.5, capacity=4096)
            d.extract_dynamic_cells(0.5, 0.25)
    d.update_measurement_grid(gen.generate_grid_host(z))
    d.update_pose(2.0, 3.0, 0.0)
    d.particle_prediction(0.1); d.get_particles()
    d.particle_assignment(); d.get_particles(); d.get_cell_ranges()
    d.grid_cell_occupancy_update(0.1); d.update_persistent_particles(); d.initialize_new_particles()
    d.statistical_moments(); d.resampling()
    d.get_grid_cells(); d.get_measurement_cells(); d.get_birth_particles(); d.get_weight_array(); d.get_joint_weight_accum()
    d.get_resampled_indices(); d.export_philox_noise(3)
    gen.close(); d.close()
print("sanitize run done")
