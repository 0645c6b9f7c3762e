// The reference's two hot-path tests (reference dogm/test/dogm_spec.cpp:10-103) restated against the drop-in C++ class
// of include/dogm/dogm.h (plain asserts: gtest is not in this image).  Built and run by tests/test_cpp_facade.py.
#include "dogm/dogm.h"
#include "mapping/laser_to_meas_grid.h"

#include <cmath>
#include <cstdio>
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

int main()
{
    if (dogm_device_count() < 1)
    {
        std::printf("no CUDA device\n");
        return 2;
    }
    int rc = test_ego_motion_compensation();
    rc |= test_predict();
    rc |= test_demo_flow();
    std::printf(rc == 0 ? "dogm_spec_b200: all passed\n" : "dogm_spec_b200: FAILED\n");
    return rc;
}
