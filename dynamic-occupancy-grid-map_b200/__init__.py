"""dogm_b200 — thin Python binding (ctypes + numpy) over the C ABI of libdogm_b200.so.

The product is the shared library (hand-written sm_100a CUDA kernels behind ``include/dogm_b200.h``); this module
only exists so that tests and ``bench.py`` can drive that ABI.  It mirrors the reference's ``dogm::DOGM`` class
(reference ``dogm/include/dogm/dogm.h:20-199``): same method names in snake_case, same argument meaning, and the
same tolerant error behaviour (CUDA errors are printed as ``GPU Kernel Error: ...`` by the library; here they
additionally raise ``DogmError`` so that tests fail loudly).

There is no CPU fallback: if the library is missing or no CUDA device is present, construction raises.

The directory name contains hyphens, so load it with ``tests/_loader.py`` (``load_dogm_b200()``), which registers
it as the module ``dogm_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DOGM_B200_LIB") or os.path.join(_HERE, "libdogm_b200.so")  # override: A/B builds of the same ABI

# ----------------------------------------------------------------------------------------------------------
# layouts (bit-identical to the reference's PODs, dogm_types.h:13-41 and dogm.h:26-60)
# ----------------------------------------------------------------------------------------------------------
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
DYNAMIC_CELL_DTYPE = np.dtype(
    [
        ("cell_idx", "<i4"),
        ("occupancy", "<f4"),
        ("mean_x_vel", "<f4"),
        ("mean_y_vel", "<f4"),
        ("var_x_vel", "<f4"),
        ("var_y_vel", "<f4"),
        ("covar_xy_vel", "<f4"),
        ("mahalanobis", "<f4"),
    ]
)
assert GRID_CELL_DTYPE.itemsize == 64 and MEAS_CELL_DTYPE.itemsize == 16 and DYNAMIC_CELL_DTYPE.itemsize == 32

RESAMPLE_SYSTEMATIC, RESAMPLE_STRATIFIED, RESAMPLE_INJECTED = 0, 1, 2
NOISE_PHILOX, NOISE_INJECTED = 0, 1


class Params(C.Structure):
    """== dogm::DOGM::Params (dogm.h:26-60)."""

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


GridParams = Params  # the name used by BASELINE.json's north_star


class BandConfig(C.Structure):
    """dogm_band_config (include/dogm_b200.h, band mode)"""

    _fields_ = [
        ("row0", C.c_int),
        ("rows", C.c_int),
        ("particle_capacity", C.c_int),
        ("birth_capacity", C.c_int),
        ("exchange_capacity", C.c_int),
        ("halo_rows", C.c_int),
        ("seed_salt", C.c_uint64),
    ]


class LaserSensorParams(C.Structure):
    """== LaserMeasurementGrid::Params (laser_to_meas_grid.h:16-22)."""

    _fields_ = [("max_range", C.c_float), ("resolution", C.c_float), ("fov", C.c_float), ("stddev_range", C.c_float)]


class Options(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("resample_mode", C.c_int), ("noise_mode", C.c_int)]


class DevicePtrs(C.Structure):
    _fields_ = [
        ("grid_cell_array", C.c_void_p),
        ("particle_array", C.c_void_p),
        ("particle_array_next", C.c_void_p),
        ("birth_particle_array", C.c_void_p),
        ("meas_cell_array", C.c_void_p),
        ("weight_array", C.c_void_p),
        ("born_masses_array", C.c_void_p),
        ("resampled_idx_array", C.c_void_p),
        ("joint_weight_accum", C.c_void_p),
        ("cell_start_array", C.c_void_p),
        ("cell_end_array", C.c_void_p),
    ]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("total_ms", C.c_double), ("launches", C.c_uint64), ("algorithmic_bytes", C.c_double)]


class BandCycleInfo(C.Structure):
    _fields_ = [("born_total", C.c_double), ("weight_total", C.c_double), ("migrated", C.c_longlong), ("phase_ms", C.c_float * 5)]


class DogmError(RuntimeError):
    pass


# every symbol include/dogm_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SYMBOLS = [
    ("dogm_create", C.c_int, [C.POINTER(Params), C.POINTER(_P)]),
    ("dogm_destroy", None, [_P]),
    ("dogm_update_grid", C.c_int, [_P, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]),
    ("dogm_update_grid_async", C.c_int, [_P, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]),
    ("dogm_synchronize", C.c_int, [_P]),
    ("dogm_update_measurement_grid", C.c_int, [_P, _P, C.c_int]),
    ("dogm_update_pose", C.c_int, [_P, C.c_float, C.c_float, C.c_float]),
    ("dogm_initialize_particles", C.c_int, [_P]),
    ("dogm_particle_prediction", C.c_int, [_P, C.c_float]),
    ("dogm_particle_assignment", C.c_int, [_P]),
    ("dogm_grid_cell_occupancy_update", C.c_int, [_P, C.c_float]),
    ("dogm_update_persistent_particles", C.c_int, [_P]),
    ("dogm_initialize_new_particles", C.c_int, [_P]),
    ("dogm_statistical_moments", C.c_int, [_P]),
    ("dogm_resampling", C.c_int, [_P]),
    ("dogm_get_grid_cells", C.c_int, [_P, _P]),
    ("dogm_get_grid_cells_begin", C.c_int, [_P, _P]),
    ("dogm_get_grid_cells_wait", C.c_int, [_P]),
    ("dogm_get_measurement_cells", C.c_int, [_P, _P]),
    ("dogm_get_particles", C.c_int, [_P, _P]),
    ("dogm_get_grid_size", C.c_int, [_P]),
    ("dogm_get_grid_cell_count", C.c_int, [_P]),
    ("dogm_get_particle_count", C.c_int, [_P]),
    ("dogm_get_new_born_particle_count", C.c_int, [_P]),
    ("dogm_get_resolution", C.c_float, [_P]),
    ("dogm_get_position_x", C.c_float, [_P]),
    ("dogm_get_position_y", C.c_float, [_P]),
    ("dogm_get_yaw", C.c_float, [_P]),
    ("dogm_get_device_ptrs", C.c_int, [_P, C.POINTER(DevicePtrs)]),
    ("dogm_set_options", C.c_int, [_P, C.POINTER(Options)]),
    ("dogm_get_options", C.c_int, [_P, C.POINTER(Options)]),
    ("dogm_set_noise", C.c_int, [_P, _P, _P, _P, _P, C.c_int]),
    ("dogm_export_philox_noise", C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    ("dogm_get_cycle_counter", C.c_uint32, [_P]),
    ("dogm_set_particles", C.c_int, [_P, _P, C.c_int]),
    ("dogm_set_birth_particles", C.c_int, [_P, _P, C.c_int]),
    ("dogm_set_grid_cells", C.c_int, [_P, _P, C.c_int]),
    ("dogm_set_measurement_cells", C.c_int, [_P, _P, C.c_int]),
    ("dogm_set_pose", C.c_int, [_P, C.c_float, C.c_float, C.c_float]),
    ("dogm_get_birth_particles", C.c_int, [_P, _P]),
    ("dogm_get_weight_array", C.c_int, [_P, _P]),
    ("dogm_get_born_masses", C.c_int, [_P, _P]),
    ("dogm_get_resampled_indices", C.c_int, [_P, _P]),
    ("dogm_get_joint_weight_accum", C.c_int, [_P, _P]),
    ("dogm_get_cell_ranges", C.c_int, [_P, _P, _P]),
    ("dogm_search_ancestors_f32", C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    ("dogm_meas_create", C.c_int, [C.POINTER(LaserSensorParams), C.c_float, C.c_float, C.POINTER(_P)]),
    ("dogm_meas_destroy", None, [_P]),
    ("dogm_meas_generate", C.c_int, [_P, _P, C.c_int, C.POINTER(_P)]),
    ("dogm_meas_generate_into", C.c_int, [_P, _P, _P, C.c_int]),
    ("dogm_meas_polar_grid", C.c_int, [_P, _P, C.c_int, _P]),
    ("dogm_meas_generate_fused", C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(_P), _P]),
    ("dogm_meas_get_grid_size", C.c_int, [_P]),
    ("dogm_extract_dynamic_cells", C.c_int, [_P, C.c_float, C.c_float, _P, C.c_int, C.POINTER(C.c_int)]),
    ("dogm_set_dynamic_cell_filter", C.c_int, [_P, C.c_float, C.c_float, C.c_int]),
    ("dogm_get_stream", _P, [_P]),
    ("dogm_get_launch_count", C.c_uint64, [_P]),
    ("dogm_kernel_timing_enable", C.c_int, [_P, C.c_int]),
    ("dogm_trace_arm", C.c_int, [_P, C.c_int]),
    ("dogm_debug_read", C.c_int, [_P, C.c_char_p, _P, C.c_size_t]),
    ("dogm_trace_read", C.c_int, [_P, C.POINTER(C.c_uint64), C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    ("dogm_kernel_timing_read", C.c_int, [_P, C.POINTER(KernelTime), C.c_int, C.POINTER(C.c_int)]),
    ("dogm_timer_start", C.c_int, [_P]),
    ("dogm_timer_stop", C.c_int, [_P]),
    ("dogm_timer_elapsed_ms", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("dogm_host_alloc_pinned", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("dogm_host_free_pinned", C.c_int, [_P]),
    ("dogm_device_alloc", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("dogm_device_free", C.c_int, [_P]),
    ("dogm_memcpy_h2d", C.c_int, [_P, _P, C.c_size_t]),
    ("dogm_memcpy_d2h", C.c_int, [_P, _P, C.c_size_t]),
    ("dogm_memcpy_d2d", C.c_int, [_P, _P, C.c_size_t]),
    ("dogm_create_band", C.c_int, [C.POINTER(Params), C.POINTER(BandConfig), C.POINTER(_P)]),
    ("dogm_band_buffer", _P, [_P, C.c_int]),
    ("dogm_band_counts", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("dogm_band_init_masses", C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_double)]),
    ("dogm_band_init_particles", C.c_int, [_P, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    ("dogm_band_predict", C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("dogm_band_append", C.c_int, [_P, C.c_int, C.c_int]),
    ("dogm_band_receive", C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, _P]),
    ("dogm_enable_peer_access", C.c_int, [C.c_int, C.c_int]),
    ("dogm_band_group_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    ("dogm_band_group_destroy", None, [_P]),
    ("dogm_band_group_set_mode", C.c_int, [_P, C.c_int]),
    ("dogm_band_set_state", C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float]),
    ("dogm_band_group_mark_initialized", C.c_int, [_P]),
    ("dogm_band_group_get_mode", C.c_int, [_P]),
    ("dogm_band_mailbox", C.c_void_p, [_P]),
    ("dogm_band_set_profile", C.c_int, [_P, C.c_int]),
    ("dogm_band_stage_times", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("dogm_band_link", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("dogm_band_cycle_enqueue", C.c_int, [_P, C.c_int, _P, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P]),
    ("dogm_band_cycle_finish", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]),
    ("dogm_band_group_band_times", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("dogm_band_group_update", C.c_int, [_P, C.POINTER(C.c_void_p), C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int),
                                        C.c_void_p]),
    ("dogm_band_update", C.c_int, [_P, _P, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_double)]),
    ("dogm_band_birth", C.c_int, [_P, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    ("dogm_band_resample", C.c_int, [_P, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    ("dogm_band_get_particles", C.c_int, [_P, _P, _P, _P, _P]),
    ("dogm_band_slot_range", C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("dogm_band_output_range", C.c_int, [C.c_uint64, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_double, C.c_double,
                                         C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    ("dogm_device_count", C.c_int, []),
    ("dogm_set_device", C.c_int, [C.c_int]),
    ("dogm_b200_version", C.c_char_p, []),
]
EXPORTED_SYMBOLS = [s[0] for s in _SYMBOLS]

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libdogm_b200.so in-tree for sm_100a (nvcc; cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise DogmError("building libdogm_b200.so failed")
    return LIB_PATH


def load_library() -> C.CDLL:
    """dlopen the C-ABI library and bind every declared symbol.  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DogmError(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or __graft_entry__.build()); there is no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in _SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _check(code: int, what: str) -> None:
    if code != 0:
        raise DogmError(f"{what} failed with code {code}")


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


class ParticlesSoA:
    """Host view of a particle block in the reference's layout (ParticlesSoA, dogm_types.h:51-144)."""

    def __init__(self, n: int, block: np.ndarray | None = None):
        self.size = n
        self.block = np.zeros(n * 28, dtype=np.uint8) if block is None else block
        assert self.block.dtype == np.uint8 and self.block.size == n * 28
        b = self.block
        self.state = b[: 16 * n].view("<f4").reshape(n, 4)
        self.grid_cell_idx = b[16 * n : 20 * n].view("<i4")
        self.weight = b[20 * n : 24 * n].view("<f4")
        self.associated = b[24 * n : 25 * n].view(np.uint8)

    @classmethod
    def from_arrays(cls, state, grid_cell_idx, weight, associated=None) -> "ParticlesSoA":
        n = len(weight)
        p = cls(n)
        p.state[:] = np.asarray(state, dtype=np.float32).reshape(n, 4)
        p.grid_cell_idx[:] = np.asarray(grid_cell_idx, dtype=np.int32)
        p.weight[:] = np.asarray(weight, dtype=np.float32)
        if associated is not None:
            p.associated[:] = np.asarray(associated, dtype=np.uint8)
        return p


class DOGM:
    """Mirror of dogm::DOGM (dogm.h:20-199) over the C ABI."""

    def __init__(self, params: Params):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.params = params
        _check(self._lib.dogm_create(C.byref(params), C.byref(self._h)), "dogm_create")
        self.grid_size = self._lib.dogm_get_grid_size(self._h)
        self.grid_cell_count = self._lib.dogm_get_grid_cell_count(self._h)
        self.particle_count = self._lib.dogm_get_particle_count(self._h)
        self.new_born_particle_count = self._lib.dogm_get_new_born_particle_count(self._h)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dogm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- the cycle (dogm.cu:115-131) -----------------------------------------------------------------------
    def update_grid(self, measurement_grid, new_x, new_y, new_yaw, dt, device=True, sync=True) -> None:
        """updateGrid(measurement_grid, new_x, new_y, new_yaw, dt, device).  `measurement_grid` is a device pointer
        (int) when device=True, a MEAS_CELL_DTYPE array when device=False, or None (keep the previous grid)."""
        if measurement_grid is not None and not device:
            assert measurement_grid.dtype == MEAS_CELL_DTYPE and measurement_grid.size == self.grid_cell_count
        fn = self._lib.dogm_update_grid if sync else self._lib.dogm_update_grid_async
        _check(fn(self._h, _ptr(measurement_grid), new_x, new_y, new_yaw, dt, 1 if device else 0), "dogm_update_grid")

    def synchronize(self) -> None:
        _check(self._lib.dogm_synchronize(self._h), "dogm_synchronize")

    def update_measurement_grid(self, measurement_grid, device=False):
        _check(
            self._lib.dogm_update_measurement_grid(self._h, _ptr(measurement_grid), 1 if device else 0),
            "dogm_update_measurement_grid",
        )

    def update_pose(self, new_x, new_y, new_yaw):
        _check(self._lib.dogm_update_pose(self._h, new_x, new_y, new_yaw), "dogm_update_pose")

    def initialize_particles(self):
        _check(self._lib.dogm_initialize_particles(self._h), "dogm_initialize_particles")

    def particle_prediction(self, dt):
        _check(self._lib.dogm_particle_prediction(self._h, dt), "dogm_particle_prediction")

    def particle_assignment(self):
        _check(self._lib.dogm_particle_assignment(self._h), "dogm_particle_assignment")

    def grid_cell_occupancy_update(self, dt):
        _check(self._lib.dogm_grid_cell_occupancy_update(self._h, dt), "dogm_grid_cell_occupancy_update")

    def update_persistent_particles(self):
        _check(self._lib.dogm_update_persistent_particles(self._h), "dogm_update_persistent_particles")

    def initialize_new_particles(self):
        _check(self._lib.dogm_initialize_new_particles(self._h), "dogm_initialize_new_particles")

    def statistical_moments(self):
        _check(self._lib.dogm_statistical_moments(self._h), "dogm_statistical_moments")

    def resampling(self):
        _check(self._lib.dogm_resampling(self._h), "dogm_resampling")

    # --- read-out (dogm.cu:133-159) ------------------------------------------------------------------------
    def get_grid_cells(self, out: np.ndarray | None = None) -> np.ndarray:
        out = np.empty(self.grid_cell_count, dtype=GRID_CELL_DTYPE) if out is None else out
        _check(self._lib.dogm_get_grid_cells(self._h, _ptr(out)), "dogm_get_grid_cells")
        return out

    def get_grid_cells_begin(self, out: np.ndarray) -> None:
        """pipelined getGridCells: starts the copy of the last cycle's cells into `out` (pinned) on the copy stream"""
        _check(self._lib.dogm_get_grid_cells_begin(self._h, _ptr(out)), "dogm_get_grid_cells_begin")

    def get_grid_cells_wait(self) -> None:
        _check(self._lib.dogm_get_grid_cells_wait(self._h), "dogm_get_grid_cells_wait")

    def get_measurement_cells(self) -> np.ndarray:
        out = np.empty(self.grid_cell_count, dtype=MEAS_CELL_DTYPE)
        _check(self._lib.dogm_get_measurement_cells(self._h, _ptr(out)), "dogm_get_measurement_cells")
        return out

    def get_particles(self) -> ParticlesSoA:
        p = ParticlesSoA(self.particle_count)
        _check(self._lib.dogm_get_particles(self._h, _ptr(p.block)), "dogm_get_particles")
        return p

    def get_birth_particles(self) -> ParticlesSoA:
        p = ParticlesSoA(self.new_born_particle_count)
        _check(self._lib.dogm_get_birth_particles(self._h, _ptr(p.block)), "dogm_get_birth_particles")
        return p

    def get_weight_array(self) -> np.ndarray:
        out = np.empty(self.particle_count, dtype=np.float32)
        _check(self._lib.dogm_get_weight_array(self._h, _ptr(out)), "dogm_get_weight_array")
        return out

    def get_born_masses(self) -> np.ndarray:
        out = np.empty(self.grid_cell_count, dtype=np.float32)
        _check(self._lib.dogm_get_born_masses(self._h, _ptr(out)), "dogm_get_born_masses")
        return out

    def get_resampled_indices(self) -> np.ndarray:
        out = np.empty(self.particle_count, dtype=np.int32)
        _check(self._lib.dogm_get_resampled_indices(self._h, _ptr(out)), "dogm_get_resampled_indices")
        return out

    def get_joint_weight_accum(self) -> np.ndarray:
        out = np.empty(self.particle_count + self.new_born_particle_count, dtype=np.float64)
        _check(self._lib.dogm_get_joint_weight_accum(self._h, _ptr(out)), "dogm_get_joint_weight_accum")
        return out

    def get_cell_ranges(self):
        s = np.empty(self.grid_cell_count, dtype=np.int32)
        e = np.empty(self.grid_cell_count, dtype=np.int32)
        _check(self._lib.dogm_get_cell_ranges(self._h, _ptr(s), _ptr(e)), "dogm_get_cell_ranges")
        return s, e

    def get_grid_size(self) -> int:
        return self._lib.dogm_get_grid_size(self._h)

    def get_resolution(self) -> float:
        return self._lib.dogm_get_resolution(self._h)

    def get_position_x(self) -> float:
        return self._lib.dogm_get_position_x(self._h)

    def get_position_y(self) -> float:
        return self._lib.dogm_get_position_y(self._h)

    def get_yaw(self) -> float:
        return self._lib.dogm_get_yaw(self._h)

    def device_ptrs(self) -> DevicePtrs:
        d = DevicePtrs()
        _check(self._lib.dogm_get_device_ptrs(self._h, C.byref(d)), "dogm_get_device_ptrs")
        return d

    # --- options / parity hooks ----------------------------------------------------------------------------
    def set_options(self, seed=123456, resample_mode=RESAMPLE_SYSTEMATIC, noise_mode=NOISE_PHILOX) -> None:
        o = Options(seed, resample_mode, noise_mode)
        _check(self._lib.dogm_set_options(self._h, C.byref(o)), "dogm_set_options")

    def set_noise(self, predict_noise=None, birth_noise=None, init_velocity=None, resample_u=None) -> None:
        def prep(a, count):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.size == count, (a.size, count)
            return a

        N, B = self.particle_count, self.new_born_particle_count
        bufs = [prep(predict_noise, 4 * N), prep(birth_noise, 2 * B), prep(init_velocity, 2 * N), prep(resample_u, N)]
        _check(self._lib.dogm_set_noise(self._h, *[_ptr(b) for b in bufs], 0), "dogm_set_noise")

    def export_philox_noise(self, cycle: int):
        N, B = self.particle_count, self.new_born_particle_count
        pn = np.empty((N, 4), np.float32)
        bn = np.empty((B, 2), np.float32)
        iv = np.empty((N, 2), np.float32)
        ru = np.empty(N, np.float32)
        _check(
            self._lib.dogm_export_philox_noise(self._h, cycle, _ptr(pn), _ptr(bn), _ptr(iv), _ptr(ru)),
            "dogm_export_philox_noise",
        )
        return pn, bn, iv, ru

    def cycle_counter(self) -> int:
        return self._lib.dogm_get_cycle_counter(self._h)

    def set_particles(self, p: ParticlesSoA) -> None:
        assert p.size == self.particle_count
        _check(self._lib.dogm_set_particles(self._h, _ptr(p.block), 0), "dogm_set_particles")

    def set_birth_particles(self, p: ParticlesSoA) -> None:
        assert p.size == self.new_born_particle_count
        _check(self._lib.dogm_set_birth_particles(self._h, _ptr(p.block), 0), "dogm_set_birth_particles")

    def set_grid_cells(self, cells: np.ndarray) -> None:
        assert cells.dtype == GRID_CELL_DTYPE and cells.size == self.grid_cell_count
        _check(self._lib.dogm_set_grid_cells(self._h, _ptr(np.ascontiguousarray(cells)), 0), "dogm_set_grid_cells")

    def set_measurement_cells(self, cells: np.ndarray) -> None:
        assert cells.dtype == MEAS_CELL_DTYPE and cells.size == self.grid_cell_count
        _check(
            self._lib.dogm_set_measurement_cells(self._h, _ptr(np.ascontiguousarray(cells)), 0),
            "dogm_set_measurement_cells",
        )

    def set_pose(self, x, y, yaw=0.0) -> None:
        _check(self._lib.dogm_set_pose(self._h, x, y, yaw), "dogm_set_pose")

    def search_ancestors_f32(self, cdf: np.ndarray, sorted_draws: np.ndarray) -> np.ndarray:
        cdf = np.ascontiguousarray(cdf, dtype=np.float32)
        draws = np.ascontiguousarray(sorted_draws, dtype=np.float32)
        out = np.empty(draws.size, dtype=np.int32)
        _check(
            self._lib.dogm_search_ancestors_f32(self._h, _ptr(cdf), cdf.size, _ptr(draws), draws.size, _ptr(out)),
            "dogm_search_ancestors_f32",
        )
        return out

    def set_dynamic_cell_filter(self, min_occupancy: float, min_velocity: float, capacity: int = 1 << 16) -> None:
        _check(
            self._lib.dogm_set_dynamic_cell_filter(self._h, min_occupancy, min_velocity, capacity),
            "dogm_set_dynamic_cell_filter",
        )

    def extract_dynamic_cells(self, min_occupancy: float, min_velocity: float, capacity: int = 1 << 16):
        out = np.empty(capacity, dtype=DYNAMIC_CELL_DTYPE)
        count = C.c_int(0)
        _check(
            self._lib.dogm_extract_dynamic_cells(self._h, min_occupancy, min_velocity, _ptr(out), capacity, C.byref(count)),
            "dogm_extract_dynamic_cells",
        )
        return out[: min(count.value, capacity)], count.value

    # --- instrumentation -----------------------------------------------------------------------------------
    def stream(self) -> int:
        return int(self._lib.dogm_get_stream(self._h) or 0)

    def launch_count(self) -> int:
        return int(self._lib.dogm_get_launch_count(self._h))

    def timer_start(self) -> None:
        _check(self._lib.dogm_timer_start(self._h), "dogm_timer_start")

    def timer_stop(self) -> None:
        _check(self._lib.dogm_timer_stop(self._h), "dogm_timer_stop")

    def timer_elapsed_ms(self) -> float:
        ms = C.c_float(0)
        _check(self._lib.dogm_timer_elapsed_ms(self._h, C.byref(ms)), "dogm_timer_elapsed_ms")
        return float(ms.value)

    def debug_read(self, name: str, out: np.ndarray) -> np.ndarray:
        _check(self._lib.dogm_debug_read(self._h, name.encode(), _ptr(out), out.nbytes), "dogm_debug_read")
        return out

    def trace_arm(self, enable: bool = True) -> None:
        _check(self._lib.dogm_trace_arm(self._h, 1 if enable else 0), "dogm_trace_arm")

    def trace_read(self) -> list:
        """[(kernel name, start stamp in ns)] of the launches since the last read, in time order."""
        cap = 64
        stamps = (C.c_uint64 * cap)()
        names = C.create_string_buffer(48 * cap)
        n = C.c_int(0)
        _check(self._lib.dogm_trace_read(self._h, stamps, names, cap, C.byref(n)), "dogm_trace_read")
        out = [(names.raw[48 * i : 48 * i + 48].split(b"\0")[0].decode(), int(stamps[i])) for i in range(n.value)]
        return sorted(out, key=lambda kv: kv[1])

    def kernel_timing_enable(self, enable: bool) -> None:
        _check(self._lib.dogm_kernel_timing_enable(self._h, 1 if enable else 0), "dogm_kernel_timing_enable")

    def kernel_timing_read(self) -> dict:
        arr = (KernelTime * 48)()
        n = C.c_int(0)
        _check(self._lib.dogm_kernel_timing_read(self._h, arr, 48, C.byref(n)), "dogm_kernel_timing_read")
        return {
            arr[i].name.decode(): {
                "total_ms": arr[i].total_ms,
                "launches": int(arr[i].launches),
                "algorithmic_bytes": arr[i].algorithmic_bytes,
            }
            for i in range(n.value)
        }


class LaserMeasurementGrid:
    """Mirror of LaserMeasurementGrid (laser_to_meas_grid.h:13-35): host beam ranges -> device MeasurementCell[]."""

    def __init__(self, params: LaserSensorParams, grid_length: float, resolution: float):
        self._lib = load_library()
        self._m = C.c_void_p()
        self.params = params
        _check(self._lib.dogm_meas_create(C.byref(params), grid_length, resolution, C.byref(self._m)), "dogm_meas_create")
        self.grid_size = self._lib.dogm_meas_get_grid_size(self._m)

    def close(self) -> None:
        if getattr(self, "_m", None) is not None and self._m:
            self._lib.dogm_meas_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def generate_grid(self, measurements) -> int:
        """generateGrid(measurements) -> device pointer owned by the generator (laser_to_meas_grid.cu:25-70)."""
        beams = np.ascontiguousarray(measurements, dtype=np.float32)
        out = C.c_void_p()
        _check(self._lib.dogm_meas_generate(self._m, _ptr(beams), beams.size, C.byref(out)), "dogm_meas_generate")
        return int(out.value)

    def generate_grid_into(self, dogm: DOGM, measurements) -> None:
        beams = np.ascontiguousarray(measurements, dtype=np.float32)
        _check(self._lib.dogm_meas_generate_into(self._m, dogm._h, _ptr(beams), beams.size), "dogm_meas_generate_into")

    def generate_grid_host(self, measurements) -> np.ndarray:
        ptr = self.generate_grid(measurements)
        out = np.empty(self.grid_size * self.grid_size, dtype=MEAS_CELL_DTYPE)
        _check(self._lib.dogm_memcpy_d2h(_ptr(out), C.c_void_p(ptr), out.nbytes), "dogm_memcpy_d2h")
        return out

    def generate_grid_fused(self, scans, want_polar: bool = False):
        """Several scans [S, K] fused in the polar grid; returns the device pointer of the measurement grid (and the fused
        polar grid [H, K, 2] when asked)."""
        scans = np.ascontiguousarray(scans, dtype=np.float32)
        assert scans.ndim == 2
        H = int(np.float32(self.params.max_range) / np.float32(self.params.resolution))
        polar = np.empty((H, scans.shape[1], 2), dtype=np.float32) if want_polar else None
        out = C.c_void_p()
        _check(
            self._lib.dogm_meas_generate_fused(self._m, _ptr(scans), scans.shape[0], scans.shape[1], C.byref(out), _ptr(polar)),
            "dogm_meas_generate_fused",
        )
        return (out.value, polar) if want_polar else out.value

    def polar_grid(self, measurements) -> np.ndarray:
        beams = np.ascontiguousarray(measurements, dtype=np.float32)
        H = int(np.float32(self.params.max_range) / np.float32(self.params.resolution))
        out = np.empty((H, beams.size, 2), dtype=np.float32)
        _check(self._lib.dogm_meas_polar_grid(self._m, _ptr(beams), beams.size, _ptr(out)), "dogm_meas_polar_grid")
        return out



# ----------------------------------------------------------------------------------------------------------
# band mode: one grid over several handles (one per band of rows; one GPU or several GPUs of this process)
# ----------------------------------------------------------------------------------------------------------
BAND_SEND_LO, BAND_SEND_HI, BAND_RECV_LO, BAND_RECV_HI, BAND_HALO_LO, BAND_HALO_HI, BAND_EDGE_LO, BAND_EDGE_HI = range(8)


def balanced_rows(row_load, n_bands: int, min_rows: int = 64):
    """Splits the rows of a grid into n_bands bands of about equal load.  row_load[y] is a non-negative estimate of the
    work in row y, e.g. the occupied mass of a measurement grid per row plus a floor for the cell work."""
    load = np.asarray(row_load, np.float64)
    G = load.size
    cum = np.concatenate([[0.0], np.cumsum(load)])
    cuts = [0]
    for r in range(1, n_bands):
        target = cum[-1] * r / n_bands
        cut = int(np.searchsorted(cum, target))
        cut = max(cut, cuts[-1] + min_rows)
        cut = min(cut, G - (n_bands - r) * min_rows)
        cuts.append(cut)
    cuts.append(G)
    return [cuts[i + 1] - cuts[i] for i in range(n_bands)]


def balanced_rows_by_phase(particles_per_row, n_bands: int, cost_particle_phases: float = 43.0, cost_particle_update: float = 37.0,
                           cost_cell_update=25.0, min_rows: int = 64):
    """Band edges for a cycle that runs in phases with a barrier after each: the particle phases (prediction, birth + CDF,
    resampling) cost a band `cost_particle_phases` per particle, the update phase (sort, per-cell sums, cell kernel)
    `cost_particle_update` per particle plus `cost_cell_update` per cell (picoseconds, measured on one B200; only the ratios
    matter).  A cycle takes  max_b(particle phases) + max_b(update phase):  the edges minimise that sum, which one load figure per
    row (balanced_rows) cannot - a band of many empty rows is cheap in particles and expensive in cells.
    For every bound P on the particle phases the smallest feasible bound U on the update phase is found by bisection; feasibility
    is a greedy sweep (a band takes rows while both bounds hold), exact for contiguous bands.
    `cost_cell_update` may be an array with one cost per row (per cell of that row): empty regions of a large grid cost the
    cell kernel a fraction of what occupied ones do (its quiet-block shortcut), which a placement measured per band can tell.
    The device-paced cycle (dogm_band_group) only meets ALL bands at the two normaliser exchanges; the prediction is coupled
    to the neighbours alone, so it belongs to the update phase there: DEVICE_PACED_COSTS."""
    n_row = np.asarray(particles_per_row, np.float64)
    G = n_row.size
    cum_n = np.concatenate([[0.0], np.cumsum(n_row)])
    cell_rows = np.broadcast_to(np.asarray(cost_cell_update, np.float64), (G,)) * G
    cum_u = cost_particle_update * cum_n + np.concatenate([[0.0], np.cumsum(cell_rows)])
    total_n, total_u = cum_n[-1], cum_u[-1]

    def sweep(P, U):
        """greedy cuts under the bounds, or None"""
        cuts = [0]
        for r in range(n_bands):
            lo = cuts[-1]
            left = n_bands - r - 1  # bands still to come: each needs min_rows rows
            end_p = int(np.searchsorted(cum_n, cum_n[lo] + P, side="right")) - 1
            end_u = int(np.searchsorted(cum_u, cum_u[lo] + U, side="right")) - 1
            end = min(end_p, end_u, G - left * min_rows)
            if end < lo + min_rows:
                if lo + min_rows > G - left * min_rows:
                    return None
                # the minimum band violates a bound: infeasible for these bounds
                return None
            cuts.append(end)
        return cuts if cuts[-1] == G else None

    best = None
    p_lo = total_n / n_bands
    for P in np.linspace(p_lo, max(p_lo * 4.0, p_lo + 1.0), 97):
        lo_u, hi_u = total_u / n_bands, total_u
        if sweep(P, hi_u) is None:
            continue
        for _ in range(40):
            mid = 0.5 * (lo_u + hi_u)
            if sweep(P, mid) is None:
                lo_u = mid
            else:
                hi_u = mid
        cuts = sweep(P, hi_u)
        n_b = np.diff(cum_n[cuts])
        u_b = np.diff(cum_u[cuts])
        t = cost_particle_phases * n_b.max() + u_b.max()
        if best is None or t < best[0]:
            best = (t, cuts)
    if best is None:
        return balanced_rows(n_row + 1.0, n_bands, min_rows)
    cuts = best[1]
    return [cuts[i + 1] - cuts[i] for i in range(n_bands)]


# picoseconds per particle in (birth + CDF + resampling), per particle in (prediction + sort + sums), per cell in the cell kernel
DEVICE_PACED_COSTS = dict(cost_particle_phases=30.0, cost_particle_update=49.0, cost_cell_update=25.0)


def band_costs_from_measurement(particles_per_band, rows, band_ms, G, cost_cell_default: float = 25.0):
    """Costs for balanced_rows_by_phase from the bands' own stage times of a profiled placement (band_ms[b] = [predict, waits,
    pull + sort + cells, birth + CDF, resample] in ms): the per-particle costs from the band with the most particles (its cell
    part taken at the default cost), and one cell cost per band - what its update stage took beyond its particles - spread over
    its rows.  Returns (cost_particle_phases, cost_particle_update, cost per row [G]) in ps, or None."""
    ms = np.asarray(band_ms, np.float64)
    n = np.asarray(particles_per_band, np.float64)
    rows = np.asarray(rows, np.int64)
    cells = rows.astype(np.float64) * G
    if ms.ndim != 2 or ms.shape[0] != n.size or n.max() <= 0:
        return None
    ps = 1e9  # ms -> ps
    b = int(np.argmax(n))
    c_phase = float(np.median(((ms[:, 3] + ms[:, 4]) * ps / np.maximum(n, 1.0))[n > 0.25 * n.max()]))
    a_b = (ms[:, 0] + ms[:, 2]) * ps
    c_upd = (a_b[b] - cost_cell_default * cells[b]) / n[b]
    if not (c_phase > 0 and c_upd > 0):
        return None
    c_cell_band = np.clip((a_b - c_upd * n) / np.maximum(cells, 1.0), 0.08 * cost_cell_default, 4.0 * cost_cell_default)
    return c_phase, float(c_upd), np.repeat(c_cell_band, rows)


def paired_band_plan(particles_per_row, n_gpus: int, cost_particle_phases: float = 30.0, cost_particle_update: float = 49.0,
                     cost_cell_update=25.0, min_rows: int = 64):
    """Two bands per GPU: 2 * n_gpus contiguous bands, band b and band 2 * n_gpus - 1 - b on the same GPU.  A scene seen by one
    sensor is dense near it and empty far away; with one contiguous band per GPU a band is either heavy in particles (it sets the
    pace of birth + CDF and of the resampling) or heavy in cells (it sets the pace of the cell kernel) and every stage waits for
    another band.  A GPU that owns one band from either end gets its share of both.  The cuts minimise
    max_gpu(particle phases) + max_gpu(update phase) for the given costs (descent over the cuts, steps from G / 16 down to one
    row).  Returns (rows per band, device per band, modelled cycle time)."""
    n_row = np.asarray(particles_per_row, np.float64)
    G = n_row.size
    R, B = n_gpus, 2 * n_gpus
    if G < B * min_rows:
        raise ValueError("grid too small for two bands per GPU")
    cum_n = np.concatenate([[0.0], np.cumsum(n_row)])
    cell_rows = np.broadcast_to(np.asarray(cost_cell_update, np.float64), (G,)) * G
    cum_c = np.concatenate([[0.0], np.cumsum(cell_rows)])
    cum_u = cost_particle_update * cum_n + cum_c
    devices = [b if b < R else B - 1 - b for b in range(B)]

    def model(cuts, tau=0.0):
        n_b = np.diff(cum_n[cuts])
        u_b = np.diff(cum_u[cuts])
        p_g = cost_particle_phases * (n_b[:R] + n_b[::-1][:R])
        u_g = u_b[:R] + u_b[::-1][:R]
        if tau <= 0.0:
            return p_g.max() + u_g.max()
        # smooth maximum: a descent over the cuts does not stall on the plateaus of the hard one
        return (p_g.max() + tau * np.log(np.exp((p_g - p_g.max()) / tau).sum())) + (u_g.max() + tau * np.log(np.exp((u_g - u_g.max()) / tau).sum()))

    # start: equal update cost per band, minimum heights enforced from both ends
    cuts = [0]
    for b in range(1, B):
        c = int(np.searchsorted(cum_u, cum_u[-1] * b / B))
        cuts.append(min(max(c, cuts[-1] + min_rows), G - (B - b) * min_rows))
    cuts.append(G)
    cuts = np.asarray(cuts, np.int64)
    scale = (cost_particle_phases * cum_n[-1] + cum_u[-1]) / R
    for tau in (0.05 * scale, 0.02 * scale, 0.005 * scale, 0.001 * scale, 0.0):
        best = model(cuts, tau)
        step = max(1, G // 16)
        while step >= 1:
            improved = True
            while improved:
                improved = False
                for i in range(1, B):
                    for d in (-step, step):
                        c = cuts[i] + d
                        if c - cuts[i - 1] < min_rows or cuts[i + 1] - c < min_rows:
                            continue
                        trial = cuts.copy()
                        trial[i] = c
                        t = model(trial, tau)
                        if t < best * (1.0 - 1e-9):
                            cuts, best, improved = trial, t, True
            step //= 2
    return [int(v) for v in np.diff(cuts)], devices, float(model(cuts))


def band_cycle_model(particles_per_row, rows, cost_particle_phases: float = 43.0, cost_particle_update: float = 37.0,
                     cost_cell_update: float = 25.0) -> float:
    """the modelled cycle time (same units as the costs) of a given split: max over bands per phase, summed"""
    n_row = np.asarray(particles_per_row, np.float64)
    G = n_row.size
    cuts = np.concatenate([[0], np.cumsum(rows)]).astype(int)
    n_b = np.array([n_row[cuts[i]:cuts[i + 1]].sum() for i in range(len(rows))])
    u_b = cost_particle_update * n_b + cost_cell_update * G * np.asarray(rows, np.float64)
    return float(cost_particle_phases * n_b.max() + u_b.max())


class BandedDOGM:
    """The orchestrator of include/dogm_b200.h's band mode for the bands of ONE process: it drives the phases of a cycle on
    every band handle and moves, between the phases, the particles that crossed a band edge (SEND -> RECV boxes), the
    halo rows of an ego-motion shift and the prefix sums of the two global normalisers (born mass, joint weight).
    `devices[r]` puts band r on that GPU: the bands then work concurrently (one host thread per band, the C calls
    release the GIL) and the copies between bands are NVLink peer copies.  By default all bands share the current GPU."""

    def __init__(self, params: Params, n_bands: int, devices=None, seed: int = 123456, slack: float = 1.75, halo_rows: int = 64,
                 resample_mode: int = RESAMPLE_SYSTEMATIC, rows=None, native: bool = True, device_paced=None):
        """rows: rows per band (default: equal shares; see balanced_rows for a split by expected particle load)
        device_paced (native orchestrator only): True / False selects how the group drives a cycle (include/dogm_b200.h,
        DOGM_BAND_GROUP_*); None keeps the library's choice (device-paced when every band's GPU reaches all the others)"""
        self._lib = load_library()
        self.params = params
        self.G = int(np.float32(params.size) / np.float32(params.resolution))
        self.R = n_bands
        base, extra = divmod(self.G, n_bands)
        self.rows = list(rows) if rows is not None else [base + (1 if r < extra else 0) for r in range(n_bands)]
        assert len(self.rows) == n_bands and sum(self.rows) == self.G and min(self.rows) > 0
        self.row0 = [sum(self.rows[:r]) for r in range(n_bands)]
        self.devices = list(devices) if devices is not None else [None] * n_bands
        n, b = params.particle_count, params.new_born_particle_count
        self.n_cap = n if n_bands == 1 else min(n, int(n / n_bands * slack) + 8192)
        self.b_cap = b if n_bands == 1 else min(b, int(b / n_bands * slack * 2) + 8192)
        self.x_cap = 0 if n_bands == 1 else max(8192, self.n_cap // 4)
        self.halo_rows = 0 if n_bands == 1 else min(halo_rows, min(self.rows))
        self.pool = None
        if devices is not None and len(set(self.devices)) > 1:
            from concurrent.futures import ThreadPoolExecutor

            self.pool = ThreadPoolExecutor(max_workers=n_bands)
        self.h = [None] * n_bands

        def create(r):
            cfg = BandConfig(self.row0[r], self.rows[r], self.n_cap, self.b_cap, self.x_cap, self.halo_rows,
                             (r * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
            hp = C.c_void_p()
            _check(self._lib.dogm_create_band(C.byref(params), C.byref(cfg), C.byref(hp)), "dogm_create_band")
            opts = Options(seed, resample_mode, NOISE_PHILOX)
            _check(self._lib.dogm_set_options(hp, C.byref(opts)), "dogm_set_options")
            self.h[r] = hp

        self._each(create)
        # neighbouring bands on different GPUs talk over NVLink: without peer access every copy would be staged through the host
        self.peer_access = []
        for r in range(n_bands - 1):
            a, b = self.devices[r], self.devices[r + 1]
            if a is not None and b is not None and a != b:
                ok = self._lib.dogm_enable_peer_access(a, b) == 0 and self._lib.dogm_enable_peer_access(b, a) == 0
                self.peer_access.append(bool(ok))
        self.first = True
        self.last_counts = []
        self.last_migration = ([], [])
        self.last_totals = {}
        # the native orchestrator (one host thread per band inside the library); `native=False` keeps the phases in Python
        self.group = None
        if native:
            arr = (C.c_void_p * n_bands)(*[h.value for h in self.h])
            g = C.c_void_p()
            _check(self._lib.dogm_band_group_create(arr, n_bands, C.byref(g)), "dogm_band_group_create")
            self.group = g
            if device_paced is not None:
                _check(self._lib.dogm_band_group_set_mode(g, 1 if device_paced else 0), "dogm_band_group_set_mode")
            self.device_paced = self._lib.dogm_band_group_get_mode(g) == 1

    def _on(self, r):
        if self.devices[r] is not None:
            set_device(self.devices[r])

    def _each(self, fn):
        """fn(r) for every band, concurrently when the bands live on different GPUs; returns the results in band order"""

        def run(r):
            self._on(r)  # the current device is a per-thread setting
            return fn(r)

        if self.pool is not None:
            return list(self.pool.map(run, range(self.R)))
        return [run(r) for r in range(self.R)]

    def close(self):
        if getattr(self, "group", None) is not None:
            self._lib.dogm_band_group_destroy(self.group)
            self.group = None
        if self.h and any(self.h):
            self._each(lambda r: self._lib.dogm_destroy(self.h[r]) if self.h[r] else None)
        self.h = []
        if self.pool is not None:
            self.pool.shutdown()
            self.pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buf(self, r, which):
        return self._lib.dogm_band_buffer(self.h[r], which)

    @staticmethod
    def _prefix(values):
        """sequential running sum (the bands must see bit-identical prefixes: before[r + 1] = before[r] + values[r])"""
        before, acc = [], 0.0
        for v in values:
            before.append(acc)
            acc = acc + float(v)
        return before, acc

    def update_grid(self, meas_band_ptrs, new_x, new_y, new_yaw, dt):
        """meas_band_ptrs[r]: device address (on band r's GPU) of the rows of the measurement grid that band r owns"""
        lib, R = self._lib, self.R
        if self.group is not None:
            ptrs = (C.c_void_p * R)(*[int(p) for p in meas_band_ptrs])
            counts = (C.c_int * R)()
            info = BandCycleInfo()
            _check(lib.dogm_band_group_update(self.group, ptrs, new_x, new_y, new_yaw, dt, counts, C.addressof(info)),
                   "dogm_band_group_update")
            self.first = False
            self.last_phase_ms = [float(v) for v in info.phase_ms]
            self.last_counts = [int(c) for c in counts]
            self.last_migration = ([int(info.migrated)], [0])
            self.last_totals = {"born": info.born_total, "weight": info.weight_total}
            times = (C.c_float * (5 * R))()
            lib.dogm_band_group_band_times(self.group, times)
            self.last_band_ms = [[float(times[r * 5 + k]) for k in range(5)] for r in range(R)]
            return self.last_counts
        if self.first:
            def masses(r):
                m = C.c_double(0.0)
                _check(lib.dogm_band_init_masses(self.h[r], C.c_void_p(meas_band_ptrs[r]), 1, C.byref(m)), "dogm_band_init_masses")
                return m.value

            before, total = self._prefix(self._each(masses))
            self._each(lambda r: _check(lib.dogm_band_init_particles(self.h[r], before[r], total, None), "dogm_band_init_particles"))
            self.first = False

        def predict(r):
            a, b = C.c_int(0), C.c_int(0)
            _check(lib.dogm_band_predict(self.h[r], new_x, new_y, new_yaw, dt, C.byref(a), C.byref(b)), "dogm_band_predict")
            return a.value, b.value

        import time as _time

        t_phase = [_time.perf_counter()]
        sent = self._each(predict)
        t_phase.append(_time.perf_counter())
        lo, hi = [s[0] for s in sent], [s[1] for s in sent]

        def receive(r):  # band r takes what its neighbours sent towards it, and their edge rows of the PREVIOUS free masses
            n_lo = hi[r - 1] if r > 0 else 0
            n_hi = lo[r + 1] if r + 1 < R else 0
            box_lo = self._buf(r - 1, BAND_SEND_HI) if n_lo else None
            box_hi = self._buf(r + 1, BAND_SEND_LO) if n_hi else None
            edge_lo = self._buf(r - 1, BAND_EDGE_HI) if (self.halo_rows and r > 0) else None
            edge_hi = self._buf(r + 1, BAND_EDGE_LO) if (self.halo_rows and r + 1 < R) else None
            self._on(r)
            _check(lib.dogm_band_receive(self.h[r], box_lo, n_lo, box_hi, n_hi, edge_lo, edge_hi), "dogm_band_receive")

        self._each(receive)  # (all bands finish this before any of them updates its cells: _each is a barrier)
        t_phase.append(_time.perf_counter())
        self.last_migration = (lo, hi)

        def update(r):
            v = C.c_double(0.0)
            _check(lib.dogm_band_update(self.h[r], C.c_void_p(meas_band_ptrs[r]), 1, dt, 1 if self.halo_rows else 0, C.byref(v)),
                   "dogm_band_update")
            return v.value

        born = self._each(update)
        t_phase.append(_time.perf_counter())
        before_b, total_b = self._prefix(born)

        def birth(r):
            v = C.c_double(0.0)
            _check(lib.dogm_band_birth(self.h[r], before_b[r], total_b, C.byref(v)), "dogm_band_birth")
            return v.value

        weight = self._each(birth)
        t_phase.append(_time.perf_counter())
        before_w, total_w = self._prefix(weight)

        def resample(r):
            n = C.c_int(0)
            _check(lib.dogm_band_resample(self.h[r], before_w[r], total_w, C.byref(n)), "dogm_band_resample")
            return n.value

        counts = self._each(resample)
        t_phase.append(_time.perf_counter())
        # host wall clock per phase [ms]: predict, exchange + append, update, birth + CDF, resample
        self.last_phase_ms = [1e3 * (b - a) for a, b in zip(t_phase[:-1], t_phase[1:])]
        self.last_counts = counts
        self.last_totals = {"born": total_b, "weight": total_w}
        return counts

    def set_state(self, state, weight, associated, grid_cells, x, y, yaw):
        """Parity hook: distributes a given particle set [n, 4] (global coordinates) and the grid cells of the whole grid over the
        bands instead of running the first cycle's initialisation."""
        state = np.ascontiguousarray(state, np.float32)
        rows_of = np.clip(state[:, 1].astype(np.int64), 0, self.G - 1)
        for r in range(self.R):
            self._on(r)
            sel = (rows_of >= self.row0[r]) & (rows_of < self.row0[r] + self.rows[r])
            st = np.ascontiguousarray(state[sel])
            w = np.ascontiguousarray(np.asarray(weight, np.float32)[sel])
            a = np.ascontiguousarray(np.asarray(associated, np.uint8)[sel])
            cells = np.ascontiguousarray(grid_cells[self.row0[r] * self.G:(self.row0[r] + self.rows[r]) * self.G])
            _check(self._lib.dogm_band_set_state(self.h[r], len(st), _ptr(st), _ptr(w), _ptr(a), _ptr(cells), x, y, yaw),
                   "dogm_band_set_state")
        self.first = False
        if self.group is not None:
            _check(self._lib.dogm_band_group_mark_initialized(self.group), "dogm_band_group_mark_initialized")

    def set_profile(self, enable: bool):
        """device-paced cycles: record events at the stage boundaries of every band (last_band_ms then holds device times)"""
        for r in range(self.R):
            self._on(r)
            _check(self._lib.dogm_band_set_profile(self.h[r], 1 if enable else 0), "dogm_band_set_profile")

    def band_handle(self, r):
        return self.h[r]

    def get_particles(self, r):
        """(state [n, 4], cell index inside the band [n], weight [n], associated [n]) of band r"""
        self._on(r)
        n = C.c_int(0)
        _check(self._lib.dogm_band_counts(self.h[r], C.byref(n), None), "dogm_band_counts")
        state = np.empty((n.value, 4), np.float32)
        idx = np.empty(n.value, np.int32)
        w = np.empty(n.value, np.float32)
        a = np.empty(n.value, np.uint8)
        _check(self._lib.dogm_band_get_particles(self.h[r], _ptr(state), _ptr(idx), _ptr(w), _ptr(a)), "dogm_band_get_particles")
        return state, idx, w, a

    def get_grid_cells(self):
        """the grid cells of all bands, stacked in row order (start_idx / end_idx refer to the band's own particle list)"""
        parts = []
        for r in range(self.R):
            self._on(r)
            out = np.empty(self.rows[r] * self.G, dtype=GRID_CELL_DTYPE)
            _check(self._lib.dogm_get_grid_cells(self.h[r], _ptr(out)), "dogm_get_grid_cells")
            parts.append(out)
        return np.concatenate(parts)


def device_count() -> int:
    return load_library().dogm_device_count()


def set_device(i: int) -> None:
    _check(load_library().dogm_set_device(i), "dogm_set_device")


def device_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    _check(load_library().dogm_device_alloc(C.byref(p), nbytes), "dogm_device_alloc")
    return int(p.value)


def device_free(ptr: int) -> None:
    _check(load_library().dogm_device_free(C.c_void_p(ptr)), "dogm_device_free")


def memcpy_h2d(dst: int, src: np.ndarray) -> None:
    _check(load_library().dogm_memcpy_h2d(C.c_void_p(dst), _ptr(np.ascontiguousarray(src)), src.nbytes), "dogm_memcpy_h2d")


def memcpy_d2h(dst: np.ndarray, src: int) -> None:
    _check(load_library().dogm_memcpy_d2h(_ptr(dst), C.c_void_p(src), dst.nbytes), "dogm_memcpy_d2h")


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by cudaMallocHost memory (never freed explicitly: lives for the process)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _check(load_library().dogm_host_alloc_pinned(C.byref(p), max(n, 16)), "dogm_host_alloc_pinned")
    buf = (C.c_uint8 * max(n, 16)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
