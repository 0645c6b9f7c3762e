// C ABI of libdogm_b200_tools.so (include/dogm_b200_tools.h): thin wrappers over the header-only simulator / DBSCAN /
// evaluator classes.  Plain C++14, no CUDA.  The symbol prefix and the class headers are taken from the include path, so a
// checker can build the very same wrappers over another implementation of these classes (DOGM_TOOLS_PREFIX).
#ifdef DOGM_TOOLS_OPEN_PRIVATE // only for builds over classes that do not expose their counters
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>
#define private public
#define protected public
#endif
#include "metrics.h"
#include "precision_evaluator.h"
#undef private
#undef protected

#include "dbscan.h"
#include "dogm_b200_tools.h"
#include "metrics.h"
#include "simulator.h"

#include <algorithm>
#include <cmath>
#include <memory>
#include <new>
#include <vector>

#ifndef DOGM_TOOLS_PREFIX
#define DOGM_TOOLS_PREFIX dogm_tools_
#endif
#define DOGM_TOOLS_CAT2(a, b) a##b
#define DOGM_TOOLS_CAT(a, b) DOGM_TOOLS_CAT2(a, b)
#define TOOL(name) DOGM_TOOLS_CAT(DOGM_TOOLS_PREFIX, name)

#ifdef DOGM_TOOLS_VEC2
using tools_vec2 = DOGM_TOOLS_VEC2;
#else
using tools_vec2 = dogm::vec2;
#endif

static Simulator make_simulator(int num_points, float fov, float grid_size, float ego_vx, float ego_vy,
                                const dogm_sim_vehicle* vehicles, int n_vehicles)
{
    Simulator sim(num_points, fov, grid_size, tools_vec2(ego_vx, ego_vy));
    for (int i = 0; i < n_vehicles; i++)
        sim.addVehicle(Vehicle(vehicles[i].width, tools_vec2(vehicles[i].x, vehicles[i].y), tools_vec2(vehicles[i].vx, vehicles[i].vy)));
    return sim;
}

struct dogm_tools_eval
{
    std::unique_ptr<PrecisionEvaluator> evaluator;
    Metric* mae;
    Metric* rmse;
    int detections;
    int steps;
};

extern "C" int TOOL(simulate)(int num_points, float fov, float grid_size, float ego_vx, float ego_vy,
                              const dogm_sim_vehicle* vehicles, int n_vehicles, int steps, float dt, float* measurements,
                              float* vehicle_states, float* ego_pose)
{
    if (num_points <= 0 || n_vehicles < 0 || steps < 0 || (n_vehicles > 0 && !vehicles) || !measurements)
        return DOGM_ERR_INVALID_ARGUMENT;
    Simulator sim = make_simulator(num_points, fov, grid_size, ego_vx, ego_vy, vehicles, n_vehicles);
    const SimulationData data = sim.update(steps, dt);
    for (int s = 0; s < steps; s++)
    {
        const SimulationStep& st = data[static_cast<std::size_t>(s)];
        for (int k = 0; k < num_points; k++)
            measurements[static_cast<std::size_t>(s) * num_points + k] = st.measurements[static_cast<std::size_t>(k)];
        if (vehicle_states)
            for (int v = 0; v < n_vehicles; v++)
            {
                float* o = vehicle_states + (static_cast<std::size_t>(s) * n_vehicles + v) * 4;
                o[0] = st.vehicles[static_cast<std::size_t>(v)].pos[0];
                o[1] = st.vehicles[static_cast<std::size_t>(v)].pos[1];
                o[2] = st.vehicles[static_cast<std::size_t>(v)].vel[0];
                o[3] = st.vehicles[static_cast<std::size_t>(v)].vel[1];
            }
        if (ego_pose)
        {
            ego_pose[2 * s] = st.ego_pose[0];
            ego_pose[2 * s + 1] = st.ego_pose[1];
        }
    }
    return 0;
}

extern "C" int TOOL(facing_side)(const dogm_sim_vehicle* vehicle, float resolution, float* out_xy, int capacity)
{
    if (!vehicle || !(resolution > 0.0f) || capacity < 0 || (capacity > 0 && !out_xy))
        return DOGM_ERR_INVALID_ARGUMENT;
    const Vehicle v(vehicle->width, tools_vec2(vehicle->x, vehicle->y), tools_vec2(vehicle->vx, vehicle->vy));
    const auto pts = v.getPointsOnFacingSide(resolution);
    for (std::size_t i = 0; i < pts.size() && static_cast<int>(i) < capacity; i++)
    {
        out_xy[2 * i] = pts[i][0];
        out_xy[2 * i + 1] = pts[i][1];
    }
    return static_cast<int>(pts.size());
}

extern "C" int TOOL(dbscan)(const float* xy, int n, float eps, int min_cells, int* out_label, int* out_clusters)
{
    if (n < 0 || (n > 0 && (!xy || !out_label)) || !out_clusters)
        return DOGM_ERR_INVALID_ARGUMENT;
    std::vector<Point<dogm::GridCell>> pts(static_cast<std::size_t>(n));
    for (int i = 0; i < n; i++)
    {
        pts[static_cast<std::size_t>(i)].x = xy[2 * i];
        pts[static_cast<std::size_t>(i)].y = xy[2 * i + 1];
        pts[static_cast<std::size_t>(i)].data = dogm::GridCell();
        pts[static_cast<std::size_t>(i)].data.start_idx = i; // the input position travels with the point
        pts[static_cast<std::size_t>(i)].cluster_id = UNCLASSIFIED;
    }
    const DBSCAN<dogm::GridCell> dbscan(eps, min_cells);
    const Clusters<dogm::GridCell> clusters = dbscan.cluster(pts);
    for (int i = 0; i < n; i++)
        out_label[i] = -1;
    for (std::size_t c = 0; c < clusters.size(); c++)
        for (const auto& p : clusters[c])
            out_label[p.data.start_idx] = static_cast<int>(c);
    *out_clusters = static_cast<int>(clusters.size());
    return 0;
}

extern "C" int TOOL(eval_create)(int num_points, float fov, float grid_size, float ego_vx, float ego_vy,
                                 const dogm_sim_vehicle* vehicles, int n_vehicles, int steps, float dt, float resolution,
                                 dogm_tools_eval** out)
{
    if (!out || num_points <= 0 || n_vehicles < 0 || steps <= 0 || (n_vehicles > 0 && !vehicles))
        return DOGM_ERR_INVALID_ARGUMENT;
    Simulator sim = make_simulator(num_points, fov, grid_size, ego_vx, ego_vy, vehicles, n_vehicles);
    dogm_tools_eval* e = new (std::nothrow) dogm_tools_eval();
    if (!e)
        return DOGM_ERR_INVALID_ARGUMENT;
    e->evaluator.reset(new PrecisionEvaluator(sim.update(steps, dt), resolution, grid_size));
    std::unique_ptr<Metric> mae(new MAE()), rmse(new RMSE());
    e->mae = mae.get();
    e->rmse = rmse.get();
    e->evaluator->registerMetric("Mean absolute error (MAE)", std::move(mae));
    e->evaluator->registerMetric("Root mean squared error (RMSE)", std::move(rmse));
    e->steps = steps;
    *out = e;
    return 0;
}

extern "C" int TOOL(eval_step)(dogm_tools_eval* e, int step, const dogm_dynamic_cell* cells, int n, int grid_size_cells)
{
    if (!e || step < 0 || step >= e->steps || n < 0 || (n > 0 && !cells) || grid_size_cells <= 0)
        return DOGM_ERR_INVALID_ARGUMENT;
    std::vector<dogm_dynamic_cell> sorted(cells, cells + n);
    std::sort(sorted.begin(), sorted.end(), [](const dogm_dynamic_cell& a, const dogm_dynamic_cell& b) { return a.cell_idx < b.cell_idx; });
    std::vector<Point<dogm::GridCell>> pts(static_cast<std::size_t>(n));
    for (int i = 0; i < n; i++)
    {
        Point<dogm::GridCell>& p = pts[static_cast<std::size_t>(i)];
        const dogm_dynamic_cell& c = sorted[static_cast<std::size_t>(i)];
        p.x = static_cast<float>(c.cell_idx % grid_size_cells);
        p.y = static_cast<float>(c.cell_idx / grid_size_cells);
        p.data = dogm::GridCell();
        p.data.mean_x_vel = c.mean_x_vel;
        p.data.mean_y_vel = c.mean_y_vel;
        p.data.var_x_vel = c.var_x_vel;
        p.data.var_y_vel = c.var_y_vel;
        p.data.covar_xy_vel = c.covar_xy_vel;
        p.cluster_id = UNCLASSIFIED;
    }
    e->evaluator->evaluateAndStoreStep(step, pts);
    return 0;
}

extern "C" int TOOL(eval_summary)(dogm_tools_eval* e, float mae[4], float rmse[4], int* detections, int* unassigned)
{
    if (!e || !mae || !rmse)
        return DOGM_ERR_INVALID_ARGUMENT;
    const PointWithVelocity a = e->mae->computeErrorStatistic(), r = e->rmse->computeErrorStatistic();
    mae[0] = a.x, mae[1] = a.y, mae[2] = a.v_x, mae[3] = a.v_y;
    rmse[0] = r.x, rmse[1] = r.y, rmse[2] = r.v_x, rmse[3] = r.v_y;
#ifdef DOGM_TOOLS_OPEN_PRIVATE
    if (unassigned)
        *unassigned = e->evaluator->number_of_unassigned_detections;
    if (detections)
        *detections = e->mae->number_of_detections;
#else
    if (unassigned)
        *unassigned = e->evaluator->unassignedDetections();
    if (detections)
        *detections = e->mae->numberOfDetections();
#endif
    return 0;
}

extern "C" void TOOL(eval_destroy)(dogm_tools_eval* e)
{
    delete e;
}
