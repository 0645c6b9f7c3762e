// precision_evaluator.h — the reference's end-to-end quality check (reference dogm/demo/utils/include/precision_evaluator.h,
// dogm/demo/utils/precision_evaluator.cpp:16-139): the dynamic cells of a scan are clustered (DBSCAN, 3 cells, 5
// neighbours), every cluster mean is matched with the nearest simulated vehicle within 5 m and the position / velocity
// errors go into the registered metrics.  Header only; `errorStatistic` / `unassignedDetections` give the numbers that the
// reference only prints.
#pragma once

#include "dbscan.h"
#include "dogm/dogm_types.h"
#include "metrics.h"
#include "simulator.h"
#include "types.h"

#include <cmath>
#include <cstddef>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

class PrecisionEvaluator
{
  public:
    PrecisionEvaluator(const SimulationData _sim_data, const float _resolution, const float _grid_size)
        : sim_data(_sim_data), resolution(_resolution), grid_size(_grid_size), number_of_unassigned_detections(0)
    {
    }

    void registerMetric(const std::string& name, std::unique_ptr<Metric> metric) { metrics.emplace(name, std::move(metric)); }

    void evaluateAndStoreStep(int simulation_step_index, const std::vector<Point<dogm::GridCell>>& cells_with_velocity,
                              bool print_current_precision = false)
    {
        const std::vector<Vehicle>& truth = sim_data[static_cast<std::size_t>(simulation_step_index)].vehicles;
        if (cells_with_velocity.empty() || truth.empty())
            return;
        const DBSCAN<dogm::GridCell> dbscan(kMaximumDbscanNeighborDistance, kMinimumNumberOfNeighbors);
        const Clusters<dogm::GridCell> clusters = dbscan.cluster(cells_with_velocity);
        int cluster_id = 0;
        for (const Cluster<dogm::GridCell>& cluster : clusters)
        {
            const PointWithVelocity mean = computeClusterMean(cluster);
            // nearest vehicle within reach; of equally near ones the first in the list (a stable order like the
            // reference's sort of the candidates by distance produces for distinct distances)
            const Vehicle* nearest = nullptr;
            float nearest_distance = 0.0f;
            for (const Vehicle& vehicle : truth)
            {
                const float distance = sqrtf(powf(mean.x - vehicle.pos[0], 2.0f) + powf(mean.y - vehicle.pos[1], 2.0f));
                if (distance < kMaximumAssignmentDistance && (!nearest || distance < nearest_distance))
                {
                    nearest = &vehicle;
                    nearest_distance = distance;
                }
            }
            if (!nearest)
            {
                ++number_of_unassigned_detections;
                continue;
            }
            PointWithVelocity current_error{};
            for (auto& metric : metrics)
                current_error = metric.second->addObjectDetection(mean, *nearest);
            if (print_current_precision)
            {
                std::cout << std::setprecision(2);
                std::cout << std::endl << "Cluster ID=" << cluster_id << std::endl;
                std::cout << "Vel. Err.: " << current_error.v_x << " " << current_error.v_y << ", Pos. Err.: " << current_error.x
                          << " " << current_error.y << std::endl;
            }
            cluster_id++;
        }
    }

    void printSummary()
    {
        for (auto& metric : metrics)
        {
            std::cout << std::endl << metric.first << ": " << std::endl;
            const PointWithVelocity error = metric.second->computeErrorStatistic();
            std::cout << "Position: " << error.x << " " << error.y << std::endl;
            std::cout << "Velocity: " << error.v_x << " " << error.v_y << std::endl;
            std::cout << std::endl;
        }
        std::cout << "Detections unassigned by evaluator: " << number_of_unassigned_detections << std::endl;
        std::cout << "Maximum possible detections: " << sim_data[0].vehicles.size() * sim_data.size() << std::endl;
    }

    // extras of this implementation
    PointWithVelocity errorStatistic(const std::string& name) { return metrics.at(name)->computeErrorStatistic(); }
    int detections(const std::string& name) { return metrics.at(name)->numberOfDetections(); }
    int unassignedDetections() const { return number_of_unassigned_detections; }

  private:
    static constexpr float kMaximumAssignmentDistance = 5.0f;
    static constexpr float kMaximumDbscanNeighborDistance = 3.0f;
    static constexpr int kMinimumNumberOfNeighbors = 5;

    // precision_evaluator.cpp:101-123: mean cell index and mean cell velocity [cells, cells/s] -> metres; the grid's y
    // axis points down from the top left corner, the world's up from the bottom left one
    PointWithVelocity computeClusterMean(const Cluster<dogm::GridCell>& cluster) const
    {
        PointWithVelocity mean;
        for (const Point<dogm::GridCell>& point : cluster)
        {
            mean.x += point.x;
            mean.y += point.y;
            mean.v_x += point.data.mean_x_vel;
            mean.v_y += point.data.mean_y_vel;
        }
        mean.x = (mean.x / cluster.size()) * resolution;
        mean.y = (mean.y / cluster.size()) * resolution;
        mean.v_x = (mean.v_x / cluster.size()) * resolution;
        mean.v_y = (mean.v_y / cluster.size()) * resolution;
        mean.v_y = -mean.v_y;
        mean.y = grid_size - mean.y;
        return mean;
    }

    SimulationData sim_data;
    float resolution;
    float grid_size;
    int number_of_unassigned_detections;
    std::map<std::string, std::unique_ptr<Metric>> metrics;
};
