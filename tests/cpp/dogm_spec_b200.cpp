// The reference's two hot-path tests (reference dogm/test/dogm_spec.cpp:10-103) restated against the drop-in C++ class
// of include/dogm/dogm.h (plain asserts: gtest is not in this image).  Built and run by tests/test_cpp_facade.py.
#include "dogm/dogm.h"
#include "mapping/laser_to_meas_grid.h"
#include "metrics.h"
#include "precision_evaluator.h"
#include "simulator.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#define CHECK(cond)                                                                                                    \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            std::printf("CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__);                                      \
            return 1;                                                                                                  \
        }                                                                                                              \
    } while (0)

static dogm::DOGM::Params spec_params()
{
    dogm::DOGM::Params p;
    p.size = 10.0f;
    p.resolution = 1.0f;
    p.particle_count = 2;
    p.new_born_particle_count = 1;
    p.persistence_prob = 0.5f;
    p.stddev_process_noise_position = 0.0f;
    p.stddev_process_noise_velocity = 0.0f;
    p.birth_prob = 0.02f;
    p.stddev_velocity = 10.0f;
    p.init_max_velocity = 30.0f;
    p.freespace_discount = 0.01f;
    return p;
}

static int test_ego_motion_compensation()
{ // dogm_spec.cpp:10-66
    dogm::DOGM dogm(spec_params());
    CHECK(dogm.last_error == 0);
    float pose_x = 10.0f, pose_y = 10.0f;
    dogm.updateGrid(nullptr, pose_x, pose_y, 0.0f, 0.0f);
    CHECK(pose_x == dogm.getPositionX());
    CHECK(pose_y == dogm.getPositionY());
    dogm::ParticlesSoA before = dogm.getParticles();
    dogm::vec4 old_state = before.state[0];
    // change lower than the resolution: no update
    dogm.updateGrid(nullptr, pose_x + 0.5f, pose_y + 0.5f, 0.0f, 0.0f);
    CHECK(pose_x == dogm.getPositionX());
    CHECK(pose_y == dogm.getPositionY());
    // +3 m: pose update and particle shift by +3 cells (moveParticlesKernel subtracts x_move = -3).  With zero
    // process noise and dt = 0 every resampled particle is a copy of a shifted one (or a newborn at cell 0's centre).
    pose_x += 3.0f;
    dogm.updateGrid(nullptr, pose_x, pose_y, 0.0f, 0.0f);
    CHECK(pose_x == dogm.getPositionX());
    CHECK(pose_y == dogm.getPositionY());
    dogm::ParticlesSoA after = dogm.getParticles();
    for (int i = 0; i < after.size; i++)
    {
        const float x = after.state[i].x;
        CHECK(x == old_state.x + 3.0f || x == 0.5f || std::fabs(x - (old_state.x + 3.0f)) < 1e-6f);
    }
    before.free();
    after.free();
    return 0;
}

static int test_predict()
{ // dogm_spec.cpp:68-103
    dogm::DOGM dogm(spec_params());
    const float dt = 0.1f;
    // give the two particles a non-trivial state first (the reference test reads uninitialised device memory)
    dogm::ParticlesSoA init(2, false);
    init.state[0] = dogm::vec4(3.25f, 4.5f, 1.7f, -2.3f);
    init.state[1] = dogm::vec4(7.125f, 1.0625f, -30.0f, 12.5f);
    init.grid_cell_idx[0] = init.grid_cell_idx[1] = 0;
    init.weight[0] = 0.37f;
    init.weight[1] = 0.63f;
    init.associated[0] = init.associated[1] = false;
    CHECK(dogm_set_particles(dogm.native(), init.memory_block, 0) == 0);

    dogm::ParticlesSoA particles = dogm.getParticles();
    dogm::vec4 old_state = particles.state[0];
    dogm::vec4 pred_state = old_state + dt * dogm::vec4(old_state[2], old_state[3], 0, 0);
    float old_weight = particles.weight[0];

    dogm.particlePrediction(dt);
    dogm::ParticlesSoA new_particles = dogm.getParticles();
    CHECK(pred_state == new_particles.state[0]);
    CHECK(old_weight * 0.5f == new_particles.weight[0]);
    init.free();
    particles.free();
    new_particles.free();
    return 0;
}

static int test_demo_flow()
{ // demo/main.cpp:21-99 in miniature: lidar scan -> measurement grid -> updateGrid -> getGridCells
    dogm::GridParams p = spec_params();
    p.size = 20.0f;
    p.resolution = 0.2f;
    p.particle_count = 20000;
    p.new_born_particle_count = 2000;
    p.persistence_prob = 0.99f;
    p.stddev_process_noise_position = 0.1f;
    p.stddev_process_noise_velocity = 1.0f;
    p.stddev_velocity = 30.0f;
    dogm::LaserSensorParams lp;
    lp.fov = 120.0f;
    lp.max_range = 20.0f;
    lp.resolution = p.resolution;
    lp.stddev_range = 0.5f;
    LaserMeasurementGrid generator(lp, p.size, p.resolution);
    dogm::DOGM grid_map(p);
    std::vector<float> scan(40, INFINITY);
    for (int i = 15; i < 25; i++)
        scan[i] = 8.0f;
    for (int step = 0; step < 5; step++)
    {
        dogm::MeasurementCell* meas = generator.generateGrid(scan);
        grid_map.updateGrid(meas, 0.0f, 0.4f * step, 0.0f, 0.1f, true);
        CHECK(grid_map.last_error == 0);
    }
    const std::vector<dogm::GridCell> cells = grid_map.getGridCells();
    CHECK((int)cells.size() == grid_map.getGridSize() * grid_map.getGridSize());
    double occ = 0.0;
    for (const auto& c : cells)
    {
        CHECK(c.occ_mass >= 0.0f && c.occ_mass <= 1.0f + 1e-5f && c.free_mass >= 0.0f);
        occ += c.occ_mass;
    }
    CHECK(occ > 1.0);
    CHECK(grid_map.getMeasurementCells().size() == cells.size());
    CHECK(grid_map.particle_array.memory_block != nullptr && grid_map.grid_cell_array != nullptr);
    return 0;
}

static int test_demo_main()
{ // demo/main.cpp:22-103 with the reference's class names: Simulator -> LaserMeasurementGrid -> DOGM -> dynamic cells ->
  // PrecisionEvaluator (MAE, RMSE); computeCellsWithVelocity (OpenCV) is replaced by DOGM::extractDynamicCells
    dogm::DOGM::Params grid_params;
    grid_params.size = 50.0f;
    grid_params.resolution = 0.2f;
    grid_params.particle_count = 3 * static_cast<int>(10e4);
    grid_params.new_born_particle_count = 3 * static_cast<int>(10e3);
    grid_params.persistence_prob = 0.99f;
    grid_params.stddev_process_noise_position = 0.1f;
    grid_params.stddev_process_noise_velocity = 1.0f;
    grid_params.birth_prob = 0.02f;
    grid_params.stddev_velocity = 30.0f;
    grid_params.init_max_velocity = 30.0f;
    grid_params.freespace_discount = 0.01f;

    LaserMeasurementGrid::Params laser_params;
    laser_params.fov = 120.0f;
    laser_params.max_range = 50.0f;
    laser_params.resolution = grid_params.resolution;
    laser_params.stddev_range = 0.5f;
    LaserMeasurementGrid grid_generator(laser_params, grid_params.size, grid_params.resolution);

    const int sensor_horizontal_scan_points = 100;
    const int num_simulation_steps = 14;
    const float simulation_step_period = 0.1f;
    const glm::vec2 ego_velocity{0.0f, 4.0f};
    const float minimum_occupancy_threshold = 0.7f;
    const float minimum_velocity_threshold = 4.0f;

    dogm::DOGM grid_map(grid_params);
    Simulator simulator(sensor_horizontal_scan_points, laser_params.fov, grid_params.size, ego_velocity);
    simulator.addVehicle(Vehicle(3.5, glm::vec2(10, 30), glm::vec2(15, 0)));
    simulator.addVehicle(Vehicle(3.0, glm::vec2(10, 20), glm::vec2(0, 5)));
    simulator.addVehicle(Vehicle(4.0, glm::vec2(35, 35), glm::vec2(0, -10)));
    simulator.addVehicle(Vehicle(1.8, glm::vec2(45, 15), glm::vec2(0, 0)));

    SimulationData sim_data = simulator.update(num_simulation_steps, simulation_step_period);
    PrecisionEvaluator precision_evaluator{sim_data, grid_params.resolution, grid_params.size};
    precision_evaluator.registerMetric("Mean absolute error (MAE)", std::unique_ptr<Metric>(new MAE()));
    precision_evaluator.registerMetric("Root mean squared error (RMSE)", std::unique_ptr<Metric>(new RMSE()));

    grid_map.setDynamicCellFilter(minimum_occupancy_threshold, minimum_velocity_threshold);
    for (int step = 0; step < num_simulation_steps; ++step)
    {
        dogm::MeasurementCell* meas_grid = grid_generator.generateGrid(sim_data[step].measurements);
        grid_map.updateGrid(meas_grid, sim_data[step].ego_pose.x, sim_data[step].ego_pose.y, 0.0f, simulation_step_period, true);
        CHECK(grid_map.last_error == 0);
        std::vector<dogm_dynamic_cell> dynamic = grid_map.extractDynamicCells(minimum_occupancy_threshold, minimum_velocity_threshold);
        std::sort(dynamic.begin(), dynamic.end(), [](const dogm_dynamic_cell& a, const dogm_dynamic_cell& b) { return a.cell_idx < b.cell_idx; });
        std::vector<Point<dogm::GridCell>> cells_with_velocity;
        for (const dogm_dynamic_cell& c : dynamic)
        {
            Point<dogm::GridCell> point;
            point.x = static_cast<float>(c.cell_idx % grid_map.getGridSize());
            point.y = static_cast<float>(c.cell_idx / grid_map.getGridSize());
            point.data = dogm::GridCell();
            point.data.mean_x_vel = c.mean_x_vel;
            point.data.mean_y_vel = c.mean_y_vel;
            point.cluster_id = UNCLASSIFIED;
            cells_with_velocity.push_back(point);
        }
        precision_evaluator.evaluateAndStoreStep(step, cells_with_velocity);
    }
    precision_evaluator.printSummary();
    const PointWithVelocity mae = precision_evaluator.errorStatistic("Mean absolute error (MAE)");
    CHECK(precision_evaluator.detections("Mean absolute error (MAE)") >= 28);
    CHECK(mae.x < 1.5f && mae.y < 1.5f && mae.v_x < 4.0f && mae.v_y < 4.0f);
    return 0;
}

static int test_streaming_loop()
{ // generateGridInto + updateGridAsync + extractDynamicCells (returns from the middle of the cycle) gives, scan
  // after scan, the list the blocking demo loop gives
    dogm::DOGM::Params grid_params;
    grid_params.size = 50.0f;
    grid_params.resolution = 0.2f;
    grid_params.particle_count = 200000;
    grid_params.new_born_particle_count = 20000;
    grid_params.persistence_prob = 0.99f;
    grid_params.stddev_process_noise_position = 0.1f;
    grid_params.stddev_process_noise_velocity = 1.0f;
    grid_params.birth_prob = 0.02f;
    grid_params.stddev_velocity = 30.0f;
    grid_params.init_max_velocity = 30.0f;
    grid_params.freespace_discount = 0.01f;
    LaserMeasurementGrid::Params laser_params;
    laser_params.fov = 120.0f;
    laser_params.max_range = 50.0f;
    laser_params.resolution = grid_params.resolution;
    laser_params.stddev_range = 0.5f;
    LaserMeasurementGrid grid_generator(laser_params, grid_params.size, grid_params.resolution);
    Simulator simulator(100, laser_params.fov, grid_params.size, glm::vec2{0.0f, 4.0f});
    simulator.addVehicle(Vehicle(3.5, glm::vec2(10, 30), glm::vec2(15, 0)));
    simulator.addVehicle(Vehicle(4.0, glm::vec2(35, 35), glm::vec2(0, -10)));
    SimulationData sim_data = simulator.update(10, 0.1f);

    dogm::DOGM blocking(grid_params), streaming(grid_params);
    blocking.setDynamicCellFilter(0.7f, 4.0f);
    streaming.setDynamicCellFilter(0.7f, 4.0f);
    size_t total = 0;
    for (int step = 0; step < 10; ++step)
    {
        dogm::MeasurementCell* meas_grid = grid_generator.generateGrid(sim_data[step].measurements);
        blocking.updateGrid(meas_grid, sim_data[step].ego_pose.x, sim_data[step].ego_pose.y, 0.0f, 0.1f, true);
        std::vector<dogm_dynamic_cell> a = blocking.extractDynamicCells(0.7f, 4.0f);

        grid_generator.generateGridInto(streaming.native(), sim_data[step].measurements);
        streaming.updateGridAsync(nullptr, sim_data[step].ego_pose.x, sim_data[step].ego_pose.y, 0.0f, 0.1f);
        CHECK(streaming.last_error == 0);
        std::vector<dogm_dynamic_cell> b = streaming.extractDynamicCells(0.7f, 4.0f);
        CHECK(streaming.last_error == 0);

        auto by_cell = [](const dogm_dynamic_cell& x, const dogm_dynamic_cell& y) { return x.cell_idx < y.cell_idx; };
        std::sort(a.begin(), a.end(), by_cell);
        std::sort(b.begin(), b.end(), by_cell);
        CHECK(a.size() == b.size());
        for (size_t k = 0; k < a.size(); k++)
            CHECK(a[k].cell_idx == b[k].cell_idx && a[k].mean_x_vel == b[k].mean_x_vel && a[k].occupancy == b[k].occupancy);
        total += a.size();
    }
    CHECK(total > 0);
    streaming.synchronize();
    const std::vector<dogm::GridCell> ga = blocking.getGridCells(), gb = streaming.getGridCells();
    CHECK(std::memcmp(ga.data(), gb.data(), ga.size() * sizeof(dogm::GridCell)) == 0);
    return 0;
}

static int test_pipelined_grid_cells()
{ // getGridCellsAsync / waitGridCells (the copy of cycle n's cells runs under cycle n + 1): the same cells as getGridCells(), the
  // public grid_cell_array follows the buffer the last cycle wrote
    dogm::GridParams p = spec_params();
    p.size = 20.0f;
    p.resolution = 0.2f;
    p.particle_count = 20000;
    p.new_born_particle_count = 2000;
    dogm::LaserSensorParams lp;
    lp.fov = 120.0f;
    lp.max_range = 20.0f;
    lp.resolution = p.resolution;
    lp.stddev_range = 0.5f;
    LaserMeasurementGrid generator(lp, p.size, p.resolution);
    dogm::DOGM blocking(p), pipelined(p);
    const size_t cells = (size_t)blocking.getGridSize() * blocking.getGridSize();
    dogm::GridCell* buf[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; k++)
        CHECK(dogm_host_alloc_pinned((void**)&buf[k], cells * sizeof(dogm::GridCell)) == 0);
    std::vector<float> scan(40, INFINITY);
    for (int i = 10; i < 30; i++)
        scan[i] = 6.0f + 0.1f * i;
    std::vector<std::vector<dogm::GridCell>> expected;
    const dogm::GridCell* seen[2] = {nullptr, nullptr};
    for (int step = 0; step < 6; step++)
    {
        dogm::MeasurementCell* meas = generator.generateGrid(scan);
        blocking.updateGrid(meas, 0.0f, 0.4f * step, 0.0f, 0.1f, true);
        expected.push_back(blocking.getGridCells());
        pipelined.updateGrid(meas, 0.0f, 0.4f * step, 0.0f, 0.1f, true);
        CHECK(pipelined.last_error == 0);
        if (step > 0)
        {
            pipelined.waitGridCells();
            CHECK(std::memcmp(buf[(step - 1) & 1], expected[step - 1].data(), cells * sizeof(dogm::GridCell)) == 0);
        }
        seen[step & 1] = pipelined.grid_cell_array;
        pipelined.getGridCellsAsync(buf[step & 1]);
    }
    pipelined.waitGridCells();
    CHECK(std::memcmp(buf[1], expected[5].data(), cells * sizeof(dogm::GridCell)) == 0);
    CHECK(seen[0] != nullptr && seen[1] != nullptr && seen[0] != seen[1]);
    return 0;
}

static int test_particles_soa_assignment()
{ // dogm_types.h:103-126: copy() and operator= copy the particles into the target's own block (any mix of host and device);
  // only an empty set adopts the other's block
    const int n = 1000;
    dogm::ParticlesSoA a(n, false), b(n, false), d1(n, true), d2(n, true);
    for (int i = 0; i < n; i++)
    {
        a.state[i] = dogm::vec4((float)i, 2.0f * i, -1.0f, 0.5f);
        a.grid_cell_idx[i] = 7 * i;
        a.weight[i] = 1.0f / (float)(i + 1);
        a.associated[i] = (i % 3) == 0;
    }
    std::memset(b.memory_block, 0, DOGM_PARTICLE_BLOCK_BYTES(n));
    d1 = a;  // host -> device
    d2 = d1; // device -> device (direct copy, no host bounce)
    b = d2;  // device -> host
    CHECK(b.memory_block != a.memory_block && d1.memory_block != d2.memory_block);
    CHECK(std::memcmp(a.memory_block, b.memory_block, DOGM_PARTICLE_BLOCK_BYTES(n)) == 0);
    dogm::ParticlesSoA c(n, false);
    c = a; // host -> host
    CHECK(c.memory_block != a.memory_block && std::memcmp(a.memory_block, c.memory_block, DOGM_PARTICLE_BLOCK_BYTES(n)) == 0);
    dogm::ParticlesSoA empty;
    empty = a; // nothing to copy into: a view
    CHECK(empty.memory_block == a.memory_block && empty.size == n && !empty.owns);
    a.free();
    b.free();
    c.free();
    d1.free();
    d2.free();
    return 0;
}

// --time: one updateGrid cycle through the C++ class (updateGrid + the refresh of its public members) against the same call
// through the C ABI, headline configuration, device-resident measurement grid; prints one JSON line
static int time_facade()
{
    dogm::DOGM::Params p;
    p.size = 120.0f;
    p.resolution = 0.1f;
    p.particle_count = 2000000;
    p.new_born_particle_count = 200000;
    p.persistence_prob = 0.99f;
    p.stddev_process_noise_position = 0.1f;
    p.stddev_process_noise_velocity = 1.0f;
    p.birth_prob = 0.02f;
    p.stddev_velocity = 30.0f;
    p.init_max_velocity = 30.0f;
    p.freespace_discount = 0.01f;
    LaserMeasurementGrid::Params lp;
    lp.fov = 120.0f;
    lp.max_range = 120.0f;
    lp.resolution = p.resolution;
    lp.stddev_range = 0.5f;
    LaserMeasurementGrid gen(lp, p.size, p.resolution);
    std::vector<float> scan(480, INFINITY);
    for (int i = 100; i < 380; i += 3)
        scan[i] = 20.0f + 0.2f * (float)(i % 50);
    dogm::MeasurementCell* meas = gen.generateGrid(scan);
    double ms[2] = {0.0, 0.0};
    const int cycles = 200;
    for (int variant = 0; variant < 2; variant++)
    {
        dogm::DOGM map(p);
        for (int k = 0; k < 10; k++)
            map.updateGrid(meas, 0.0f, 0.4f * k, 0.0f, 0.1f, true);
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 10; k < 10 + cycles; k++)
        {
            if (variant == 0)
                map.updateGrid(meas, 0.0f, 0.4f * k, 0.0f, 0.1f, true);
            else
                dogm_update_grid(map.native(), meas, 0.0f, 0.4f * k, 0.0f, 0.1f, 1);
        }
        ms[variant] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / cycles;
    }
    std::printf("{\"facade_update_grid_ms\": %.5f, \"c_abi_update_grid_ms\": %.5f, \"refresh_overhead_us\": %.2f, \"cycles\": %d}\n", ms[0],
                ms[1], (ms[0] - ms[1]) * 1e3, cycles);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc > 1 && !std::strcmp(argv[1], "--time"))
        return dogm_device_count() < 1 ? 2 : time_facade();
    if (dogm_device_count() < 1)
    {
        std::printf("no CUDA device\n");
        return 2;
    }
    int rc = test_ego_motion_compensation();
    rc |= test_predict();
    rc |= test_demo_flow();
    rc |= test_demo_main();
    rc |= test_streaming_loop();
    rc |= test_particles_soa_assignment();
    rc |= test_pipelined_grid_cells();
    std::printf(rc == 0 ? "dogm_spec_b200: all passed\n" : "dogm_spec_b200: FAILED\n");
    return rc;
}
