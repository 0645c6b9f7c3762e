"""ctypes wrapper of oracle/_ref/libdogm_ref.so — the reference's own CUDA implementation, compiled unmodified from
/root/reference by oracle/build_ref.sh together with oracle/ref_harness.cu.

TEST INFRASTRUCTURE ONLY (tests/, bench.py --impl reference).  Needs a GPU to run; `available()` is False when the
library has not been built (it is built in the authoring container and travels to the GPU box as a prebuilt file).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdogm_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build() -> bool:
    """Builds oracle/_ref when the reference sources are present; returns whether the library exists afterwards."""
    if os.path.isdir("/root/reference/dogm/src"):
        res = subprocess.run(["bash", os.path.join(_HERE, "build_ref.sh")], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + res.stdout[-3000:] + res.stderr[-3000:])
    return available()


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    lib = C.CDLL(LIB_PATH)
    VP = C.c_void_p
    lib.ref_create.restype = VP
    lib.ref_create.argtypes = [VP]
    lib.ref_destroy.argtypes = [VP]
    lib.ref_grid_size.argtypes = [VP]
    lib.ref_rng_thread_count.argtypes = [VP]
    lib.ref_update_grid.argtypes = [VP, VP, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
    lib.ref_update_grid_timed.restype = C.c_double
    lib.ref_update_grid_timed.argtypes = [VP, VP, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
    lib.ref_extract_noise.argtypes = [VP, C.c_int, VP, VP, VP, VP]
    lib.ref_stage_update_measurement_grid.argtypes = [VP, VP, C.c_int]
    lib.ref_stage_update_pose.argtypes = [VP, C.c_float, C.c_float, C.c_float]
    lib.ref_stage_particle_prediction.argtypes = [VP, C.c_float]
    lib.ref_stage_grid_cell_occupancy_update.argtypes = [VP, C.c_float]
    for name in (
        "ref_stage_particle_assignment",
        "ref_stage_update_persistent_particles",
        "ref_stage_initialize_new_particles",
        "ref_stage_statistical_moments",
        "ref_stage_publish",
    ):
        getattr(lib, name).argtypes = [VP]
    lib.ref_stage_resampling_verbose.argtypes = [VP, VP, VP, VP, C.POINTER(C.c_float)]
    for name in (
        "ref_get_particles",
        "ref_get_particles_next",
        "ref_get_birth_particles",
        "ref_set_particles",
        "ref_get_grid_cells",
        "ref_set_grid_cells",
        "ref_get_meas_cells",
        "ref_get_weight_array",
        "ref_get_born_masses",
        "ref_get_pose",
    ):
        getattr(lib, name).argtypes = [VP, VP]
    lib.ref_device_alloc.restype = VP
    lib.ref_device_alloc.argtypes = [C.c_size_t]
    lib.ref_device_free.argtypes = [VP]
    lib.ref_memcpy_h2d.argtypes = [VP, VP, C.c_size_t]
    lib.ref_polar_grid.argtypes = [VP, C.c_int, C.c_int, C.c_float, C.c_float, VP]
    if hasattr(lib, "ref_polar_fused"):
        lib.ref_polar_fused.argtypes = [VP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, VP]
    assert lib.ref_sizeof_particle() == 28 and lib.ref_sizeof_grid_cell() == 64
    _lib = lib
    return lib


class RefDOGM:
    """The reference dogm::DOGM behind the stage names of dogm_b200.DOGM.  `params` is any ctypes struct with the
    reference's Params layout; particle blocks are numpy uint8 arrays of n*28 bytes in the reference's layout."""

    def __init__(self, params, grid_cell_dtype, meas_cell_dtype):
        self._lib = load_library()
        self._params = params
        self._h = self._lib.ref_create(C.addressof(params))
        self.grid_size = self._lib.ref_grid_size(self._h)
        self.grid_cell_count = self.grid_size * self.grid_size
        self.particle_count = params.particle_count
        self.new_born_particle_count = params.new_born_particle_count
        self._gdt, self._mdt = grid_cell_dtype, meas_cell_dtype

    def close(self):
        if self._h:
            self._lib.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _p(a):
        if a is None:
            return None
        if isinstance(a, int):
            return C.c_void_p(a)
        return C.c_void_p(a.ctypes.data)

    def update_grid(self, meas, x, y, yaw, dt, device=False, timed=False):
        meas = None if meas is None else (meas if isinstance(meas, int) else np.ascontiguousarray(meas))
        if timed:
            return self._lib.ref_update_grid_timed(self._h, self._p(meas), x, y, yaw, dt, 1 if device else 0)
        self._lib.ref_update_grid(self._h, self._p(meas), x, y, yaw, dt, 1 if device else 0)

    def extract_noise(self, first_cycle: bool):
        """(init_velocity, predict_noise, birth_noise, resample_unit_sorted) of the NEXT cycle."""
        N, B = self.particle_count, self.new_born_particle_count
        iv = np.zeros((N, 2), np.float32)
        pn = np.zeros((N, 4), np.float32)
        bn = np.zeros((B, 2), np.float32)
        ru = np.zeros(N, np.float32)
        self._lib.ref_extract_noise(self._h, 1 if first_cycle else 0, self._p(iv), self._p(pn), self._p(bn), self._p(ru))
        return iv, pn, bn, np.sort(ru)

    def stage_update_measurement_grid(self, meas):
        meas = None if meas is None else np.ascontiguousarray(meas)
        self._lib.ref_stage_update_measurement_grid(self._h, self._p(meas), 0)

    def stage_update_pose(self, x, y, yaw):
        self._lib.ref_stage_update_pose(self._h, x, y, yaw)

    def stage_particle_prediction(self, dt):
        self._lib.ref_stage_particle_prediction(self._h, dt)

    def stage_particle_assignment(self):
        self._lib.ref_stage_particle_assignment(self._h)

    def stage_grid_cell_occupancy_update(self, dt):
        self._lib.ref_stage_grid_cell_occupancy_update(self._h, dt)

    def stage_update_persistent_particles(self):
        self._lib.ref_stage_update_persistent_particles(self._h)

    def stage_initialize_new_particles(self):
        self._lib.ref_stage_initialize_new_particles(self._h)

    def stage_statistical_moments(self):
        self._lib.ref_stage_statistical_moments(self._h)

    def stage_resampling_verbose(self):
        N, B = self.particle_count, self.new_born_particle_count
        cdf = np.zeros(N + B, np.float32)
        rand_sorted = np.zeros(N, np.float32)
        idx = np.zeros(N, np.int32)
        jm = C.c_float(0)
        self._lib.ref_stage_resampling_verbose(self._h, self._p(cdf), self._p(rand_sorted), self._p(idx), C.byref(jm))
        return cdf, rand_sorted, idx, float(jm.value)

    def stage_publish(self):
        self._lib.ref_stage_publish(self._h)

    def run_cycle_verbose(self, meas, x, y, yaw, dt, first_cycle: bool) -> dict:
        """One updateGrid (dogm.cu:115-131) driven stage by stage with a snapshot after every stage.  Keys:
        noise iv/pn/bn/ru; P0,G0,pose0 = state the cycle starts from (after first-cycle initialisation, before the pose
        update); P1 after prediction; P2,G2 after assignment; G3,born3,W3,P3 after the occupancy update; W4,G4 after the
        persistent-weight update; BP5,G5 after birth; G6 after the moments; cdf7,rand7,idx7,joint_max7,P7 after
        resampling; pose7."""
        s = {}
        s["iv"], s["pn"], s["bn"], s["ru"] = self.extract_noise(first_cycle)
        self.stage_update_measurement_grid(meas)
        s["P0"] = self.get_particles_block()
        s["G0"] = self.get_grid_cells()
        s["pose0"] = np.array(self.get_pose(), np.float32)
        self.stage_update_pose(x, y, yaw)
        self.stage_particle_prediction(dt)
        s["P1"] = self.get_particles_block()
        self.stage_particle_assignment()
        s["P2"] = self.get_particles_block()
        s["G2"] = self.get_grid_cells()
        self.stage_grid_cell_occupancy_update(dt)
        s["G3"] = self.get_grid_cells()
        s["born3"] = self.get_born_masses()
        s["W3"] = self.get_weight_array()
        s["P3"] = self.get_particles_block()
        self.stage_update_persistent_particles()
        s["W4"] = self.get_weight_array()
        s["G4"] = self.get_grid_cells()
        self.stage_initialize_new_particles()
        s["BP5"] = self.get_birth_particles_block()
        s["G5"] = self.get_grid_cells()
        self.stage_statistical_moments()
        s["G6"] = self.get_grid_cells()
        s["cdf7"], s["rand7"], s["idx7"], jm = self.stage_resampling_verbose()
        s["joint_max7"] = np.array([jm], np.float32)
        s["P7"] = self.get_particles_next_block()
        self.stage_publish()
        s["pose7"] = np.array(self.get_pose(), np.float32)
        return s

    def _block(self, fn, n):
        b = np.zeros(n * 28, np.uint8)
        fn(self._h, self._p(b))
        return b

    def get_particles_block(self):
        return self._block(self._lib.ref_get_particles, self.particle_count)

    def get_particles_next_block(self):
        return self._block(self._lib.ref_get_particles_next, self.particle_count)

    def get_birth_particles_block(self):
        return self._block(self._lib.ref_get_birth_particles, self.new_born_particle_count)

    def set_particles_block(self, block):
        self._lib.ref_set_particles(self._h, self._p(np.ascontiguousarray(block)))

    def get_grid_cells(self):
        out = np.zeros(self.grid_cell_count, dtype=self._gdt)
        self._lib.ref_get_grid_cells(self._h, self._p(out))
        return out

    def set_grid_cells(self, cells):
        self._lib.ref_set_grid_cells(self._h, self._p(np.ascontiguousarray(cells)))

    def get_meas_cells(self):
        out = np.zeros(self.grid_cell_count, dtype=self._mdt)
        self._lib.ref_get_meas_cells(self._h, self._p(out))
        return out

    def get_weight_array(self):
        out = np.zeros(self.particle_count, np.float32)
        self._lib.ref_get_weight_array(self._h, self._p(out))
        return out

    def get_born_masses(self):
        out = np.zeros(self.grid_cell_count, np.float32)
        self._lib.ref_get_born_masses(self._h, self._p(out))
        return out

    def get_pose(self):
        out = np.zeros(3, np.float32)
        self._lib.ref_get_pose(self._h, self._p(out))
        return tuple(float(v) for v in out)


def device_upload(arr: np.ndarray) -> int:
    lib = load_library()
    arr = np.ascontiguousarray(arr)
    p = lib.ref_device_alloc(arr.nbytes)
    lib.ref_memcpy_h2d(C.c_void_p(p), C.c_void_p(arr.ctypes.data), arr.nbytes)
    return int(p)


def device_free(ptr: int) -> None:
    load_library().ref_device_free(C.c_void_p(ptr))


def polar_fused(scans, height: int, resolution: float, stddev_range: float) -> np.ndarray:
    """the reference's polar grid after createPolarGridTextureKernel(scan 0) and fusePolarGridTextureKernel(scans 1..)"""
    lib = load_library()
    scans = np.ascontiguousarray(scans, dtype=np.float32)
    out = np.empty((height, scans.shape[1], 2), dtype=np.float32)
    rc = lib.ref_polar_fused(C.c_void_p(scans.ctypes.data), scans.shape[0], scans.shape[1], height, resolution, stddev_range,
                             C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"ref_polar_fused failed: {rc}")
    return out


def polar_grid(beams, height: int, resolution: float, stddev_range: float) -> np.ndarray:
    """createPolarGridTextureKernel (measurement_grid.cu:72-89) run on a plain CUDA surface: [height][beams][occ, free]."""
    lib = load_library()
    beams = np.ascontiguousarray(beams, dtype=np.float32)
    out = np.zeros((height, beams.size, 2), np.float32)
    rc = lib.ref_polar_grid(C.c_void_p(beams.ctypes.data), beams.size, height, resolution, stddev_range, C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"ref_polar_grid failed: {rc}")
    return out
