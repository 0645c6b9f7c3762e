"""dogm_b200.tools — ctypes binding of libdogm_b200_tools.so (include/dogm_b200_tools.h): the lidar scene simulator, DBSCAN
and the MAE / RMSE evaluator of the reference's demo as host-side companions of the DOGM path (plain C++, no CUDA).

`Tools(lib_path, prefix)` binds any library that exports the same entry points under another prefix (a checker uses this to
drive a second implementation through identical calls)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdogm_b200_tools.so")

ENTRY_POINTS = ("simulate", "facing_side", "dbscan", "eval_create", "eval_step", "eval_summary", "eval_destroy")

DYNAMIC_CELL_DTYPE = np.dtype(
    [("cell_idx", "<i4"), ("occupancy", "<f4"), ("mean_x_vel", "<f4"), ("mean_y_vel", "<f4"), ("var_x_vel", "<f4"),
     ("var_y_vel", "<f4"), ("covar_xy_vel", "<f4"), ("mahalanobis", "<f4")]
)


class SimVehicle(C.Structure):
    _fields_ = [("width", C.c_float), ("x", C.c_float), ("y", C.c_float), ("vx", C.c_float), ("vy", C.c_float)]


# the four vehicles of the reference's default scene (demo/main.cpp:70-73): (width, x, y, vx, vy)
DEMO_VEHICLES = [(3.5, 10, 30, 15, 0), (3.0, 10, 20, 0, 5), (4.0, 35, 35, 0, -10), (1.8, 45, 15, 0, 0)]
# its alternative scene (demo/main.cpp:64-66)
DEMO_VEHICLES_ALT = [(4.0, 10, 25, 10, -8), (6.0, 40, 30, -8, 6), (3.0, 48, 15, -12, 0)]


class ToolsError(RuntimeError):
    pass


def _vehicles(vehicles):
    arr = (SimVehicle * max(1, len(vehicles)))()
    for i, v in enumerate(vehicles):
        arr[i] = SimVehicle(*[float(t) for t in v])
    return arr


class Tools:
    def __init__(self, lib_path: str = LIB_PATH, prefix: str = "dogm_tools_"):
        if not os.path.exists(lib_path):
            raise ToolsError(f"{lib_path} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}`")
        self.lib = C.CDLL(lib_path)
        self.prefix = prefix
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        sig = {
            "simulate": (C.c_int, [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp, C.c_int, C.c_int, C.c_float, vp, vp, vp]),
            "facing_side": (C.c_int, [vp, C.c_float, vp, C.c_int]),
            "dbscan": (C.c_int, [vp, C.c_int, C.c_float, C.c_int, vp, ip]),
            "eval_create": (C.c_int, [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(vp)]),
            "eval_step": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int]),
            "eval_summary": (C.c_int, [vp, fp, fp, ip, ip]),
            "eval_destroy": (None, [vp]),
        }
        self.fn = {}
        for name in ENTRY_POINTS:
            f = getattr(self.lib, prefix + name)  # AttributeError if the library does not export it
            f.restype, f.argtypes = sig[name]
            self.fn[name] = f

    def _check(self, code, what):
        if code != 0:
            raise ToolsError(f"{self.prefix}{what} failed with code {code}")

    def simulate(self, num_points, fov, grid_size, ego_velocity, vehicles, steps, dt):
        """-> measurements [steps, num_points], vehicle states [steps, n, 4] = (x, y, vx, vy), ego pose [steps, 2]"""
        n = len(vehicles)
        meas = np.empty((steps, num_points), np.float32)
        states = np.empty((steps, max(n, 1), 4), np.float32)
        ego = np.empty((steps, 2), np.float32)
        self._check(
            self.fn["simulate"](num_points, fov, grid_size, ego_velocity[0], ego_velocity[1], C.cast(_vehicles(vehicles), C.c_void_p), n,
                                steps, dt, meas.ctypes.data, states.ctypes.data, ego.ctypes.data),
            "simulate",
        )
        return meas, states[:, :n], ego

    def facing_side(self, vehicle, resolution, capacity=4096):
        out = np.empty((capacity, 2), np.float32)
        v = SimVehicle(*[float(t) for t in vehicle])
        count = self.fn["facing_side"](C.cast(C.pointer(v), C.c_void_p), resolution, out.ctypes.data, capacity)
        if count < 0:
            raise ToolsError(f"{self.prefix}facing_side failed with code {count}")
        return out[: min(count, capacity)].copy(), count

    def dbscan(self, xy, eps, min_cells):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        labels = np.empty(len(xy), np.int32)
        n_clusters = C.c_int(0)
        self._check(self.fn["dbscan"](xy.ctypes.data, len(xy), eps, min_cells, labels.ctypes.data, C.byref(n_clusters)), "dbscan")
        return labels, n_clusters.value

    def evaluator(self, num_points, fov, grid_size, ego_velocity, vehicles, steps, dt, resolution):
        return Evaluator(self, num_points, fov, grid_size, ego_velocity, vehicles, steps, dt, resolution)


class Evaluator:
    def __init__(self, tools, num_points, fov, grid_size, ego_velocity, vehicles, steps, dt, resolution):
        self.t = tools
        self.h = C.c_void_p()
        tools._check(
            tools.fn["eval_create"](num_points, fov, grid_size, ego_velocity[0], ego_velocity[1], C.cast(_vehicles(vehicles), C.c_void_p),
                                    len(vehicles), steps, dt, resolution, C.byref(self.h)),
            "eval_create",
        )

    def step(self, step, cells, grid_size_cells):
        cells = np.ascontiguousarray(cells, DYNAMIC_CELL_DTYPE)
        self.t._check(self.t.fn["eval_step"](self.h, step, cells.ctypes.data, len(cells), grid_size_cells), "eval_step")

    def summary(self):
        mae, rmse = (C.c_float * 4)(), (C.c_float * 4)()
        det, un = C.c_int(0), C.c_int(0)
        self.t._check(self.t.fn["eval_summary"](self.h, mae, rmse, C.byref(det), C.byref(un)), "eval_summary")
        return {"mae": np.array(mae[:], np.float32), "rmse": np.array(rmse[:], np.float32), "detections": det.value, "unassigned": un.value}

    def close(self):
        if self.h:
            self.t.fn["eval_destroy"](self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
