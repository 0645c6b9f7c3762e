// laser_to_meas_grid.h — drop-in LaserMeasurementGrid over the C ABI (reference
// dogm/demo/simulator/include/mapping/laser_to_meas_grid.h:13-35, laser_to_meas_grid.cu:10-70).
// One CUDA kernel replaces the reference's kernel -> OpenGL render -> kernel path; no GL context is needed.
#ifndef DOGM_B200_LASER_TO_MEAS_GRID_H
#define DOGM_B200_LASER_TO_MEAS_GRID_H

#include "../dogm/dogm_types.h"

#include <vector>

class LaserMeasurementGrid
{
  public:
    using Params = ::dogm_laser_params; // max_range, resolution, fov, stddev_range (laser_to_meas_grid.h:16-22)

    LaserMeasurementGrid(const Params& params, float grid_length, float resolution)
    {
        dogm_meas_create(&params, grid_length, resolution, &handle);
    }
    ~LaserMeasurementGrid() { dogm_meas_destroy(handle); }
    LaserMeasurementGrid(const LaserMeasurementGrid&) = delete;
    LaserMeasurementGrid& operator=(const LaserMeasurementGrid&) = delete;

    // host beam ranges (inf = no return) -> device MeasurementCell[grid_size^2] owned by this object
    dogm::MeasurementCell* generateGrid(const std::vector<float>& measurements)
    {
        dogm::MeasurementCell* out = nullptr;
        dogm_meas_generate(handle, measurements.data(), static_cast<int>(measurements.size()), &out);
        return out;
    }
    // several scans of the same sensor geometry (equal sizes), fused in the polar grid (the reference's unused
    // fusePolarGridTextureKernel, measurement_grid.cu:91-113)
    dogm::MeasurementCell* generateGridFused(const std::vector<std::vector<float>>& scans)
    {
        dogm::MeasurementCell* out = nullptr;
        if (scans.empty())
            return out;
        std::vector<float> flat;
        for (const auto& s : scans)
            flat.insert(flat.end(), s.begin(), s.end());
        dogm_meas_generate_fused(handle, flat.data(), static_cast<int>(scans.size()), static_cast<int>(scans[0].size()), &out, nullptr);
        return out;
    }
    // straight into a DOGM's measurement buffer; follow with updateGrid(nullptr, ...)
    void generateGridInto(::dogm_handle* dogm, const std::vector<float>& measurements)
    {
        dogm_meas_generate_into(handle, dogm, measurements.data(), static_cast<int>(measurements.size()));
    }

  private:
    ::dogm_meas_handle* handle = nullptr;
};

namespace dogm
{
using LaserSensorParams = ::dogm_laser_params; // the name BASELINE.json's north_star uses
}

#endif
