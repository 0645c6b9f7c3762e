/*
 * dogm_b200_tools.h — C ABI of the host-side companions of the DOGM path (libdogm_b200_tools.so, plain C++, no CUDA):
 * the reference demo's lidar scene simulator (reference dogm/demo/simulator/include/simulator.h:11-62) and its evaluator
 * (DBSCAN + MAE / RMSE against the simulated vehicles, reference dogm/demo/utils/include/dbscan.h:9-48,
 * precision_evaluator.h:20-43, metrics.h:11-63).  The C++ classes themselves are header-only (include/simulator.h, dbscan.h,
 * metrics.h, precision_evaluator.h, same names as the reference's); this ABI exists for bindings and tests.
 * Every function returns 0 on success, a negative DOGM_ERR_* value otherwise, unless stated.
 */
#ifndef DOGM_B200_TOOLS_H
#define DOGM_B200_TOOLS_H

#include "dogm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Vehicle(width, position, velocity), simulator.h:16-19 */
typedef struct dogm_sim_vehicle
{
    float width, x, y, vx, vy;
} dogm_sim_vehicle;

/* Simulator(num_points, fov, grid_size, ego_velocity) + addVehicle(...) + update(steps, dt), simulator.cpp:45-83.
 * measurements[steps * num_points] (INFINITY = no return), vehicle_states[steps * n_vehicles * 4] = (x, y, vx, vy) after
 * each step, ego_pose[steps * 2]. */
int dogm_tools_simulate(int num_points, float fov, float grid_size, float ego_vx, float ego_vy,
                        const dogm_sim_vehicle* vehicles, int n_vehicles, int steps, float dt, float* measurements,
                        float* vehicle_states, float* ego_pose);

/* Vehicle::getPointsOnFacingSide(resolution), simulator.cpp:12-27; returns the number of points (at most `capacity`
 * (x, y) pairs are written) or a negative error. */
int dogm_tools_facing_side(const dogm_sim_vehicle* vehicle, float resolution, float* out_xy, int capacity);

/* DBSCAN<GridCell>(eps, min_cells).cluster(points), dbscan.cpp:25-56: out_label[i] = index of the returned cluster that
 * holds point i, -1 if none; *out_clusters = number of clusters returned (empty ones included). */
int dogm_tools_dbscan(const float* xy, int n, float eps, int min_cells, int* out_label, int* out_clusters);

/* PrecisionEvaluator over the scene simulated with the given parameters (precision_evaluator.cpp:26-30); MAE and RMSE are
 * registered as in demo/main.cpp:77-79. */
typedef struct dogm_tools_eval dogm_tools_eval;
int dogm_tools_eval_create(int num_points, float fov, float grid_size, float ego_vx, float ego_vy,
                           const dogm_sim_vehicle* vehicles, int n_vehicles, int steps, float dt, float resolution,
                           dogm_tools_eval** out);
/* evaluateAndStoreStep(step, cells), precision_evaluator.cpp:37-99; the cells are the records of
 * dogm_extract_dynamic_cells (any order: they are put into the row-major order of computeCellsWithVelocity,
 * image_creation.cpp:29-62), grid_size_cells = DOGM::getGridSize(). */
int dogm_tools_eval_step(dogm_tools_eval* e, int step, const dogm_dynamic_cell* cells, int n, int grid_size_cells);
/* the numbers printSummary prints (precision_evaluator.cpp:125-139): (x, y, v_x, v_y) errors, matched detections, and the
 * clusters no vehicle was near */
int dogm_tools_eval_summary(dogm_tools_eval* e, float mae[4], float rmse[4], int* detections, int* unassigned);
void dogm_tools_eval_destroy(dogm_tools_eval* e);

#ifdef __cplusplus
}
#endif

#endif /* DOGM_B200_TOOLS_H */
