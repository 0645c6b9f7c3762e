"""ctypes wrapper of oracle/libdogm_oracle.so — the CPU restatement of the reference DOGM cycle.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product path (dynamic-occupancy-grid-map_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdogm_oracle.so")

GRID_CELL_DTYPE = np.dtype(
    [
        ("start_idx", "<i4"),
        ("end_idx", "<i4"),
        ("new_born_occ_mass", "<f4"),
        ("pers_occ_mass", "<f4"),
        ("free_mass", "<f4"),
        ("occ_mass", "<f4"),
        ("pred_occ_mass", "<f4"),
        ("mu_A", "<f4"),
        ("mu_UA", "<f4"),
        ("w_A", "<f4"),
        ("w_UA", "<f4"),
        ("mean_x_vel", "<f4"),
        ("mean_y_vel", "<f4"),
        ("var_x_vel", "<f4"),
        ("var_y_vel", "<f4"),
        ("covar_xy_vel", "<f4"),
    ]
)
MEAS_CELL_DTYPE = np.dtype([("free_mass", "<f4"), ("occ_mass", "<f4"), ("likelihood", "<f4"), ("p_A", "<f4")])

RESAMPLE_SYSTEMATIC, RESAMPLE_STRATIFIED, RESAMPLE_INJECTED = 0, 1, 2
SUM_F64, SUM_F32_SCAN = 0, 1


class Params(C.Structure):
    _fields_ = [
        ("size", C.c_float),
        ("resolution", C.c_float),
        ("particle_count", C.c_int),
        ("new_born_particle_count", C.c_int),
        ("persistence_prob", C.c_float),
        ("stddev_process_noise_position", C.c_float),
        ("stddev_process_noise_velocity", C.c_float),
        ("birth_prob", C.c_float),
        ("stddev_velocity", C.c_float),
        ("init_max_velocity", C.c_float),
        ("freespace_discount", C.c_float),
    ]


class LaserParams(C.Structure):
    _fields_ = [("max_range", C.c_float), ("resolution", C.c_float), ("fov", C.c_float), ("stddev_range", C.c_float)]


class _Particles(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("state", C.POINTER(C.c_float)),
        ("grid_cell_idx", C.POINTER(C.c_int)),
        ("weight", C.POINTER(C.c_float)),
        ("associated", C.POINTER(C.c_uint8)),
    ]


class _Dogm(C.Structure):
    _fields_ = [
        ("params", Params),
        ("grid_size", C.c_int),
        ("grid_cell_count", C.c_int),
        ("particle_count", C.c_int),
        ("new_born_particle_count", C.c_int),
        ("grid_cell_array", C.c_void_p),
        ("meas_cell_array", C.c_void_p),
        ("particle_array", _Particles),
        ("particle_array_next", _Particles),
        ("birth_particle_array", _Particles),
        ("weight_array", C.POINTER(C.c_float)),
        ("born_masses_array", C.POINTER(C.c_float)),
        ("joint_weight_accum", C.POINTER(C.c_double)),
        ("joint_weight_accum_f32", C.POINTER(C.c_float)),
        ("resampled_idx", C.POINTER(C.c_int)),
        ("birth_slot_end", C.POINTER(C.c_int)),
        ("joint_max", C.c_float),
        ("predict_noise", C.c_void_p),
        ("birth_noise", C.c_void_p),
        ("init_velocity", C.c_void_p),
        ("resample_u", C.c_void_p),
        ("resample_mode", C.c_int),
        ("sum_mode", C.c_int),
        ("first_pose_received", C.c_int),
        ("first_measurement_received", C.c_int),
        ("position_x", C.c_float),
        ("position_y", C.c_float),
        ("yaw", C.c_float),
        ("cycle", C.c_uint32),
    ]


_lib = None


def build() -> str:
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER(_Dogm)
    VP = C.c_void_p
    lib.oracle_create.restype = P
    lib.oracle_create.argtypes = [C.POINTER(Params)]
    lib.oracle_destroy.argtypes = [P]
    lib.oracle_set_noise.argtypes = [P, VP, VP, VP, VP]
    lib.oracle_set_modes.argtypes = [P, C.c_int, C.c_int]
    lib.oracle_update_grid.argtypes = [P, VP, C.c_float, C.c_float, C.c_float, C.c_float]
    lib.oracle_update_measurement_grid.argtypes = [P, VP]
    lib.oracle_update_pose.argtypes = [P, C.c_float, C.c_float, C.c_float]
    for name in (
        "oracle_initialize_particles",
        "oracle_particle_assignment",
        "oracle_update_persistent_particles",
        "oracle_initialize_new_particles",
        "oracle_statistical_moments",
        "oracle_resampling",
    ):
        getattr(lib, name).argtypes = [P]
    lib.oracle_particle_prediction.argtypes = [P, C.c_float]
    lib.oracle_grid_cell_occupancy_update.argtypes = [P, C.c_float]
    lib.oracle_search_ancestors_f32.argtypes = [VP, C.c_int, VP, C.c_int, VP]
    lib.oracle_meas_grid_size.restype = C.c_int
    lib.oracle_meas_grid_size.argtypes = [C.c_float, C.c_float]
    lib.oracle_meas_polar_height.restype = C.c_int
    lib.oracle_meas_polar_height.argtypes = [C.POINTER(LaserParams)]
    lib.oracle_meas_polar_grid.argtypes = [C.POINTER(LaserParams), VP, C.c_int, VP]
    lib.oracle_meas_polar_fuse.argtypes = [C.POINTER(LaserParams), VP, VP, C.c_int]
    lib.oracle_meas_generate_fused.argtypes = [C.POINTER(LaserParams), C.c_float, C.c_float, VP, C.c_int, C.c_int, VP, VP]
    lib.oracle_meas_generate.argtypes = [C.POINTER(LaserParams), C.c_float, C.c_float, VP, C.c_int, VP]
    lib.oracle_extract_dynamic_cells.restype = C.c_int
    lib.oracle_extract_dynamic_cells.argtypes = [VP, C.c_int, C.c_float, C.c_float, VP, C.c_int]
    _lib = lib
    return lib


def _np(ptr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(count,)).view(dtype)


class _ParticleView:
    def __init__(self, p: _Particles):
        n = p.n
        self.size = n
        self.state = _np(p.state, 4 * n, np.float32).reshape(n, 4)
        self.grid_cell_idx = _np(p.grid_cell_idx, n, np.int32)
        self.weight = _np(p.weight, n, np.float32)
        self.associated = _np(p.associated, n, np.uint8)

    def copy(self):
        c = object.__new__(_ParticleView)
        c.size = self.size
        c.state = self.state.copy()
        c.grid_cell_idx = self.grid_cell_idx.copy()
        c.weight = self.weight.copy()
        c.associated = self.associated.copy()
        return c


class OracleDOGM:
    """The oracle behind the same method names as dogm_b200.DOGM (views alias the oracle's memory: copy to keep)."""

    def __init__(self, params: Params, resample_mode=RESAMPLE_INJECTED, sum_mode=SUM_F64):
        self._lib = load_library()
        self._o = self._lib.oracle_create(C.byref(params))
        self._keep = {}
        self._lib.oracle_set_modes(self._o, resample_mode, sum_mode)
        o = self._o.contents
        self.grid_size = o.grid_size
        self.grid_cell_count = o.grid_cell_count
        self.particle_count = o.particle_count
        self.new_born_particle_count = o.new_born_particle_count

    def close(self):
        if self._o:
            self._lib.oracle_destroy(self._o)
            self._o = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_modes(self, resample_mode, sum_mode=SUM_F64):
        self._lib.oracle_set_modes(self._o, resample_mode, sum_mode)

    def set_noise(self, predict_noise=None, birth_noise=None, init_velocity=None, resample_u=None):
        ptrs = []
        for key, a in (("p", predict_noise), ("b", birth_noise), ("i", init_velocity), ("r", resample_u)):
            if a is None:
                ptrs.append(None)
            else:
                a = np.ascontiguousarray(a, dtype=np.float32)
                self._keep[key] = a  # the oracle borrows the pointer
                ptrs.append(C.c_void_p(a.ctypes.data))
        self._lib.oracle_set_noise(self._o, *ptrs)

    # stages
    def update_grid(self, meas, x, y, yaw, dt):
        ptr = None
        if meas is not None:
            assert meas.dtype == MEAS_CELL_DTYPE and meas.size == self.grid_cell_count
            meas = np.ascontiguousarray(meas)
            ptr = C.c_void_p(meas.ctypes.data)
        self._lib.oracle_update_grid(self._o, ptr, x, y, yaw, dt)

    def update_measurement_grid(self, meas):
        ptr = None
        if meas is not None:
            meas = np.ascontiguousarray(meas)
            ptr = C.c_void_p(meas.ctypes.data)
        self._lib.oracle_update_measurement_grid(self._o, ptr)

    def update_pose(self, x, y, yaw):
        self._lib.oracle_update_pose(self._o, x, y, yaw)

    def initialize_particles(self):
        self._lib.oracle_initialize_particles(self._o)

    def particle_prediction(self, dt):
        self._lib.oracle_particle_prediction(self._o, dt)

    def particle_assignment(self):
        self._lib.oracle_particle_assignment(self._o)

    def grid_cell_occupancy_update(self, dt):
        self._lib.oracle_grid_cell_occupancy_update(self._o, dt)

    def update_persistent_particles(self):
        self._lib.oracle_update_persistent_particles(self._o)

    def initialize_new_particles(self):
        self._lib.oracle_initialize_new_particles(self._o)

    def statistical_moments(self):
        self._lib.oracle_statistical_moments(self._o)

    def resampling(self):
        self._lib.oracle_resampling(self._o)

    # views
    @property
    def grid_cells(self) -> np.ndarray:
        o = self._o.contents
        buf = (C.c_uint8 * (64 * o.grid_cell_count)).from_address(o.grid_cell_array)
        return np.frombuffer(buf, dtype=GRID_CELL_DTYPE)

    @property
    def meas_cells(self) -> np.ndarray:
        o = self._o.contents
        buf = (C.c_uint8 * (16 * o.grid_cell_count)).from_address(o.meas_cell_array)
        return np.frombuffer(buf, dtype=MEAS_CELL_DTYPE)

    @property
    def particles(self) -> _ParticleView:
        return _ParticleView(self._o.contents.particle_array)

    @property
    def particles_next(self) -> _ParticleView:
        return _ParticleView(self._o.contents.particle_array_next)

    @property
    def birth_particles(self) -> _ParticleView:
        return _ParticleView(self._o.contents.birth_particle_array)

    @property
    def weight_array(self) -> np.ndarray:
        return _np(self._o.contents.weight_array, self.particle_count, np.float32)

    @property
    def born_masses(self) -> np.ndarray:
        return _np(self._o.contents.born_masses_array, self.grid_cell_count, np.float32)

    @property
    def joint_weight_accum(self) -> np.ndarray:
        return _np(self._o.contents.joint_weight_accum, self.particle_count + self.new_born_particle_count, np.float64)

    @property
    def joint_weight_accum_f32(self) -> np.ndarray:
        return _np(self._o.contents.joint_weight_accum_f32, self.particle_count + self.new_born_particle_count, np.float32)

    @property
    def resampled_idx(self) -> np.ndarray:
        return _np(self._o.contents.resampled_idx, self.particle_count, np.int32)

    @property
    def birth_slot_end(self) -> np.ndarray:
        return _np(self._o.contents.birth_slot_end, self.grid_cell_count, np.int32)

    @property
    def joint_max(self) -> float:
        return float(self._o.contents.joint_max)

    @property
    def position(self):
        o = self._o.contents
        return float(o.position_x), float(o.position_y), float(o.yaw)

    def set_pose(self, x, y, yaw=0.0):
        o = self._o.contents
        o.position_x, o.position_y, o.yaw = x, y, yaw
        o.first_pose_received = 1

    def set_first_measurement_received(self, flag=True):
        self._o.contents.first_measurement_received = 1 if flag else 0


def search_ancestors_f32(cdf, sorted_draws) -> np.ndarray:
    lib = load_library()
    cdf = np.ascontiguousarray(cdf, dtype=np.float32)
    draws = np.ascontiguousarray(sorted_draws, dtype=np.float32)
    out = np.empty(draws.size, dtype=np.int32)
    lib.oracle_search_ancestors_f32(cdf.ctypes.data, cdf.size, draws.ctypes.data, draws.size, out.ctypes.data)
    return out


def meas_generate(laser: LaserParams, grid_length: float, resolution: float, beams) -> np.ndarray:
    lib = load_library()
    beams = np.ascontiguousarray(beams, dtype=np.float32)
    gs = lib.oracle_meas_grid_size(grid_length, resolution)
    out = np.empty(gs * gs, dtype=MEAS_CELL_DTYPE)
    lib.oracle_meas_generate(C.byref(laser), grid_length, resolution, beams.ctypes.data, beams.size, out.ctypes.data)
    return out


def meas_polar_fused(laser: LaserParams, scans) -> np.ndarray:
    """polar grid [H, K, 2] after fusing all scans [S, K] (createPolarGridTextureKernel, then fusePolarGridTextureKernel)"""
    lib = load_library()
    scans = np.ascontiguousarray(scans, np.float32)
    table = meas_polar_grid(laser, scans[0])
    for s in range(1, scans.shape[0]):
        lib.oracle_meas_polar_fuse(C.byref(laser), table.ctypes.data, scans[s].ctypes.data, scans.shape[1])
    return table


def meas_generate_fused(laser: LaserParams, grid_length: float, resolution: float, scans) -> np.ndarray:
    lib = load_library()
    scans = np.ascontiguousarray(scans, np.float32)
    H = lib.oracle_meas_polar_height(C.byref(laser))
    gs = int(np.float32(grid_length) / np.float32(resolution))
    table = np.empty((H, scans.shape[1], 2), np.float32)
    out = np.empty(gs * gs, dtype=MEAS_CELL_DTYPE)
    lib.oracle_meas_generate_fused(C.byref(laser), grid_length, resolution, scans.ctypes.data, scans.shape[0], scans.shape[1],
                                   table.ctypes.data, out.ctypes.data)
    return out


def meas_polar_grid(laser: LaserParams, beams) -> np.ndarray:
    lib = load_library()
    beams = np.ascontiguousarray(beams, dtype=np.float32)
    H = lib.oracle_meas_polar_height(C.byref(laser))
    out = np.empty((H, beams.size, 2), dtype=np.float32)
    lib.oracle_meas_polar_grid(C.byref(laser), beams.ctypes.data, beams.size, out.ctypes.data)
    return out


def extract_dynamic_cells(cells: np.ndarray, min_occupancy: float, min_velocity: float, capacity: int = 1 << 16):
    lib = load_library()
    cells = np.ascontiguousarray(cells)
    rec = np.empty((capacity, 8), dtype=np.float32)
    n = lib.oracle_extract_dynamic_cells(cells.ctypes.data, cells.size, min_occupancy, min_velocity, rec.ctypes.data, capacity)
    return rec[: min(n, capacity)], n
