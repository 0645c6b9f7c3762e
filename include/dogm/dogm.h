// dogm.h — drop-in C++ class dogm::DOGM over the C ABI of libdogm_b200.so.
//
// Mirrors the reference's public class (reference dogm/include/dogm/dogm.h:20-199): the same Params, constructor,
// updateGrid, getters, the eight public stage methods and the public data members callers use
// (demo/main.cpp:58,88-92; demo/utils/image_creation.cpp:27,70,92,114,153-155; test/dogm_spec.cpp:24-60,84-99).
// A reference user changes the include path to this directory and links libdogm_b200.so instead of the static `dogm`.
//
// Differences a caller can observe:
//   * errors: the library prints "GPU Kernel Error: ..." like the reference's CHECK_ERROR and carries on; the last
//     error code is also kept in `last_error`;
//   * particle_array_next aliases particle_array after a cycle (the reference holds two equal deep copies);
//   * the device pointers in the public members are refreshed after every call (sets swap roles inside a cycle);
//   * names used by BASELINE.json exist as aliases: dogm::GridParams, dogm::LaserSensorParams (mapping header).
#pragma once

#include "dogm_types.h"

#include <vector>

namespace dogm
{

class DOGM
{
  public:
    using Params = ::dogm_params; // dogm.h:26-60: size, resolution, particle_count, new_born_particle_count, ...

    explicit DOGM(const Params& params_) : params(params_)
    {
        last_error = dogm_create(&params, &handle);
        grid_size = dogm_get_grid_size(handle);
        grid_cell_count = dogm_get_grid_cell_count(handle);
        particle_count = dogm_get_particle_count(handle);
        new_born_particle_count = dogm_get_new_born_particle_count(handle);
        refresh();
    }
    ~DOGM() { dogm_destroy(handle); }
    DOGM(const DOGM&) = delete; // the reference class is copyable only by accident (double free)
    DOGM& operator=(const DOGM&) = delete;

    // dogm.h:82-83 / dogm.cu:115-131
    void updateGrid(MeasurementCell* measurement_grid, float new_x, float new_y, float new_yaw, float dt, bool device = true)
    {
        last_error = dogm_update_grid(handle, measurement_grid, new_x, new_y, new_yaw, dt, device ? 1 : 0);
        refresh();
    }

    std::vector<GridCell> getGridCells() const // dogm.cu:133-141
    {
        std::vector<GridCell> cells(static_cast<size_t>(grid_cell_count));
        dogm_get_grid_cells(handle, cells.data());
        return cells;
    }
    // getGridCells without stalling the filter (not in the reference): the cells of the last updateGrid start travelling into
    // `pinned_out` (cudaMallocHost memory of grid_cell_count cells) and arrive under the following updateGrid calls;
    // waitGridCells() returns when they are there.  grid_cell_array alternates between two device buffers from then on.
    void getGridCellsAsync(GridCell* pinned_out)
    {
        dogm_get_grid_cells_begin(handle, pinned_out);
    }
    void waitGridCells()
    {
        dogm_get_grid_cells_wait(handle);
    }
    std::vector<MeasurementCell> getMeasurementCells() const // dogm.cu:143-151
    {
        std::vector<MeasurementCell> cells(static_cast<size_t>(grid_cell_count));
        dogm_get_measurement_cells(handle, cells.data());
        return cells;
    }
    ParticlesSoA getParticles() const // dogm.cu:153-159: host copy, the caller frees it
    {
        ParticlesSoA particles(particle_count, false);
        dogm_get_particles(handle, particles.memory_block);
        return particles;
    }

    int getGridSize() const { return grid_size; }
    float getResolution() const { return params.resolution; }
    float getPositionX() const { return dogm_get_position_x(handle); }
    float getPositionY() const { return dogm_get_position_y(handle); }
    float getYaw() const { return dogm_get_yaw(handle); }

    // the public stage methods, dogm.h:148-156
    void initializeParticles() { stage(dogm_initialize_particles(handle)); }
    void particlePrediction(float dt) { stage(dogm_particle_prediction(handle, dt)); }
    void particleAssignment() { stage(dogm_particle_assignment(handle)); }
    void gridCellOccupancyUpdate(float dt) { stage(dogm_grid_cell_occupancy_update(handle, dt)); }
    void updatePersistentParticles() { stage(dogm_update_persistent_particles(handle)); }
    void initializeNewParticles() { stage(dogm_initialize_new_particles(handle)); }
    void statisticalMoments() { stage(dogm_statistical_moments(handle)); }
    void resampling() { stage(dogm_resampling(handle)); }

    // extras of this implementation
    // the cycle without the device synchronisation of dogm.cu:130: returns once the kernels are enqueued.  Pair it with
    // setDynamicCellFilter + extractDynamicCells (which returns as soon as the cycle has produced its list) or synchronize().
    void updateGridAsync(MeasurementCell* measurement_grid, float new_x, float new_y, float new_yaw, float dt, bool device = true)
    {
        last_error = dogm_update_grid_async(handle, measurement_grid, new_x, new_y, new_yaw, dt, device ? 1 : 0);
    }
    void synchronize() { last_error = dogm_synchronize(handle); }
    std::vector<::dogm_dynamic_cell> extractDynamicCells(float min_occupancy, float min_velocity, int capacity = 1 << 16)
    {
        std::vector<::dogm_dynamic_cell> out(static_cast<size_t>(capacity));
        int count = 0;
        last_error = dogm_extract_dynamic_cells(handle, min_occupancy, min_velocity, out.data(), capacity, &count);
        out.resize(static_cast<size_t>(count < capacity ? count : capacity));
        return out;
    }
    // registers the filter so that each updateGrid compacts the list itself (capacity 0 = off)
    void setDynamicCellFilter(float min_occupancy, float min_velocity, int capacity = 1 << 16)
    {
        last_error = dogm_set_dynamic_cell_filter(handle, min_occupancy, min_velocity, capacity);
    }
    void setOptions(const ::dogm_options& o) { last_error = dogm_set_options(handle, &o); }
    ::dogm_handle* native() { return handle; }

  public: // dogm.h:158-191
    Params params;

    GridCell* grid_cell_array = nullptr;
    ParticlesSoA particle_array;
    ParticlesSoA particle_array_next;
    ParticlesSoA birth_particle_array;
    MeasurementCell* meas_cell_array = nullptr;

    float* weight_array = nullptr;
    float* born_masses_array = nullptr;

    int grid_size = 0;
    int grid_cell_count = 0;
    int particle_count = 0;
    int new_born_particle_count = 0;

    int last_error = 0;

  private:
    void stage(int e)
    {
        last_error = e;
        refresh();
    }
    void refresh()
    {
        ::dogm_device_ptrs p;
        if (dogm_get_device_ptrs(handle, &p) != 0)
            return;
        grid_cell_array = p.grid_cell_array;
        meas_cell_array = p.meas_cell_array;
        weight_array = p.weight_array;
        born_masses_array = p.born_masses_array;
        // (rebind, not operator=: assigning a ParticlesSoA copies the particles, as in the reference)
        particle_array.rebind(p.particle_array, particle_count, true);
        particle_array_next.rebind(p.particle_array_next, particle_count, true);
        birth_particle_array.rebind(p.birth_particle_array, new_born_particle_count, true);
    }

    ::dogm_handle* handle = nullptr;
};

using GridParams = DOGM::Params; // the name BASELINE.json's north_star uses

} // namespace dogm
