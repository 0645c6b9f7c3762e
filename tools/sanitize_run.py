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
    # streaming loop: scan resampled inside the cell kernel, list published from the middle of the cycle
    for c in range(3):
        z = np.where(rng.uniform(size=37) < 0.6, rng.uniform(2, size * 0.9, 37), np.inf).astype(np.float32)
        gen.generate_grid_into(d, z)
        d.update_grid(None, 1.2 + 0.3 * c, 1.8 + 0.45 * c, 0.0, 0.1, sync=False)
        d.extract_dynamic_cells(0.6, 0.5, capacity=4096)
    d.synchronize()
    d.update_measurement_grid(gen.generate_grid_host(z))
    d.update_pose(2.0, 3.0, 0.0)
    d.particle_prediction(0.1); d.get_particles()
    d.particle_assignment(); d.get_particles(); d.get_cell_ranges()
    d.grid_cell_occupancy_update(0.1); d.update_persistent_particles(); d.initialize_new_particles()
    d.statistical_moments(); d.resampling()
    d.get_grid_cells(); d.get_measurement_cells(); d.get_birth_particles(); d.get_weight_array(); d.get_joint_weight_accum()
    d.get_resampled_indices(); d.export_philox_noise(3)
    # pipelined read-out: copy stream, two GridCell buffers in turn
    bufs = [gpu.pinned_empty((d.grid_cell_count,), gpu.GRID_CELL_DTYPE) for _ in range(2)]
    for c in range(4):
        z = np.where(rng.uniform(size=37) < 0.6, rng.uniform(2, size * 0.9, 37), np.inf).astype(np.float32)
        d.update_grid(gen.generate_grid(z), 3.0 + 0.3 * c, 4.0, 0.0, 0.1, device=True, sync=False)
        d.get_grid_cells_wait()
        d.get_grid_cells_begin(bufs[c & 1])
    d.get_grid_cells_wait()
    gen.close(); d.close()
# band mode: two and three bands on one GPU (outbox compaction, append, halo rows, global normalisers)
size, res, n, b = 24.0, 0.25, 30000, 3000
gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(size, res, 120.0, 0.5), size, res)
rng = np.random.default_rng(2)
for bands in (2, 3):
    bd = gpu.BandedDOGM(make_params(gpu, size, res, n, b), bands, seed=5, halo_rows=8)
    dev = gpu.device_alloc(bd.G * bd.G * 16)
    for c in range(4):
        z = np.where(rng.uniform(size=48) < 0.6, rng.uniform(2, size * 0.9, 48), np.inf).astype(np.float32)
        gpu.memcpy_h2d(dev, gen.generate_grid_host(z))
        bd.update_grid([dev + bd.row0[r] * bd.G * 16 for r in range(bands)], 0.2 * c, 0.6 * c, 0.0, 0.1)
    bd.get_grid_cells()
    bd.close()
    gpu.device_free(dev)
gen.close()
print("sanitize run done")
