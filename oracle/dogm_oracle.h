/*
 * dogm_oracle.h — CPU restatement of the reference DOGM update cycle.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (dynamic-occupancy-grid-map_b200/) never links, imports or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference/dogm).
 * Pinning status: see the header of dogm_oracle.c.
 */
#ifndef DOGM_ORACLE_H
#define DOGM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_params /* == DOGM::Params, include/dogm/dogm.h:26-60 */
{
    float size;
    float resolution;
    int particle_count;
    int new_born_particle_count;
    float persistence_prob;
    float stddev_process_noise_position;
    float stddev_process_noise_velocity;
    float birth_prob;
    float stddev_velocity;
    float init_max_velocity;
    float freespace_discount;
} oracle_params;

typedef struct oracle_grid_cell /* == GridCell, include/dogm/dogm_types.h:13-33 */
{
    int start_idx;
    int end_idx;
    float new_born_occ_mass;
    float pers_occ_mass;
    float free_mass;
    float occ_mass;
    float pred_occ_mass;
    float mu_A;
    float mu_UA;
    float w_A;
    float w_UA;
    float mean_x_vel;
    float mean_y_vel;
    float var_x_vel;
    float var_y_vel;
    float covar_xy_vel;
} oracle_grid_cell;

typedef struct oracle_meas_cell /* == MeasurementCell, include/dogm/dogm_types.h:35-41 */
{
    float free_mass;
    float occ_mass;
    float likelihood;
    float p_A;
} oracle_meas_cell;

/* ParticlesSoA (include/dogm/dogm_types.h:51-144) as four plain arrays */
typedef struct oracle_particles
{
    int n;
    float* state; /* n * 4: x, y, vx, vy */
    int* grid_cell_idx;
    float* weight;
    uint8_t* associated;
} oracle_particles;

#define ORACLE_RESAMPLE_SYSTEMATIC 0
#define ORACLE_RESAMPLE_STRATIFIED 1
#define ORACLE_RESAMPLE_INJECTED 2

/* Summation mode for the order-dependent reductions (per-cell weight sums, born-mass prefix, joint CDF).
 *   ORACLE_SUM_F64 : sequential double accumulation (what the new implementation is checked against bit-for-bit)
 *   ORACLE_SUM_F32_SCAN : the reference's literal scheme — sequential float inclusive scan over the whole array and
 *                         segment sums as scan differences (common.h:12-32); used to bound the reference's own error */
#define ORACLE_SUM_F64 0
#define ORACLE_SUM_F32_SCAN 1

typedef struct oracle_dogm
{
    oracle_params params;
    int grid_size;
    int grid_cell_count;
    int particle_count;
    int new_born_particle_count;

    oracle_grid_cell* grid_cell_array;
    oracle_meas_cell* meas_cell_array;
    oracle_particles particle_array;
    oracle_particles particle_array_next;
    oracle_particles birth_particle_array;
    float* weight_array;      /* N */
    float* born_masses_array; /* C */
    double* joint_weight_accum; /* N + B, running sum of [weight_array, birth weights] */
    float* joint_weight_accum_f32; /* N + B, float view (reference semantics) */
    int* resampled_idx;       /* N */
    int* birth_slot_end;      /* C, exclusive end slot of every cell of the last birth / init distribution */
    float joint_max;          /* weight total of the last resampling (float, dogm.cu:400) */

    /* injected noise (borrowed pointers, set by oracle_set_noise) */
    const float* predict_noise; /* N*4, scaled */
    const float* birth_noise;   /* B*2, scaled */
    const float* init_velocity; /* N*2 */
    const float* resample_u;    /* INJECTED: N ascending fractions; SYSTEMATIC: [0]; STRATIFIED: N uniforms */
    int resample_mode;
    int sum_mode;

    int first_pose_received;
    int first_measurement_received;
    float position_x, position_y, yaw;
    uint32_t cycle;
} oracle_dogm;

oracle_dogm* oracle_create(const oracle_params* params);
void oracle_destroy(oracle_dogm* o);
void oracle_set_noise(oracle_dogm* o, const float* predict_noise, const float* birth_noise,
                      const float* init_velocity, const float* resample_u);
void oracle_set_modes(oracle_dogm* o, int resample_mode, int sum_mode);

/* DOGM::updateGrid, src/dogm.cu:115-131 (meas may be NULL: keep the previous grid, dogm_spec.cpp:32) */
void oracle_update_grid(oracle_dogm* o, const oracle_meas_cell* meas, float new_x, float new_y, float new_yaw, float dt);

/* the stages, src/dogm.cu:161-423 */
void oracle_update_measurement_grid(oracle_dogm* o, const oracle_meas_cell* meas);
void oracle_update_pose(oracle_dogm* o, float new_x, float new_y, float new_yaw);
void oracle_initialize_particles(oracle_dogm* o);
void oracle_particle_prediction(oracle_dogm* o, float dt);
void oracle_particle_assignment(oracle_dogm* o);
void oracle_grid_cell_occupancy_update(oracle_dogm* o, float dt);
void oracle_update_persistent_particles(oracle_dogm* o);
void oracle_initialize_new_particles(oracle_dogm* o);
void oracle_statistical_moments(oracle_dogm* o);
void oracle_resampling(oracle_dogm* o);

/* thrust::lower_bound over a float CDF whose last entry was replaced by the largest draw, resampling.cu:34-47 */
void oracle_search_ancestors_f32(const float* cdf, int n_cdf, const float* sorted_draws, int n_draws, int* out);

/* LaserMeasurementGrid, demo/simulator/mapping: inverse sensor model (kernel/measurement_grid.cu:33-89) and the
 * polar->cartesian warp (opengl/renderer.cpp:11-31,50-54,66-84; opengl/shader.cpp:22-35; measurement_grid.cu:115-132) */
typedef struct oracle_laser_params
{
    float max_range;
    float resolution;
    float fov;
    float stddev_range;
} oracle_laser_params;
int oracle_meas_grid_size(float grid_length, float resolution);
int oracle_meas_polar_height(const oracle_laser_params* p);
void oracle_meas_polar_grid(const oracle_laser_params* p, const float* beams, int num_beams, float* out /* H*K*2 */);
void oracle_meas_generate(const oracle_laser_params* p, float grid_length, float resolution, const float* beams,
                          int num_beams, oracle_meas_cell* out /* gs*gs */);
/* multi-scan fusion in the polar grid (fusePolarGridTextureKernel + combine_masses, measurement_grid.cu:13-31,91-113) */
void oracle_meas_polar_fuse(const oracle_laser_params* p, float* table /* H*K*2, in place */, const float* beams, int num_beams);
void oracle_meas_generate_fused(const oracle_laser_params* p, float grid_length, float resolution, const float* scans,
                                int num_scans, int num_beams, float* table /* H*K*2 scratch */, oracle_meas_cell* out);

/* computeCellsWithVelocity, demo/utils/image_creation.cpp:19-66; returns the count, fills up to capacity records of
 * 8 floats {cell_idx (as int bits), occupancy, mean_x, mean_y, var_x, var_y, covar, mahalanobis} */
int oracle_extract_dynamic_cells(const oracle_grid_cell* cells, int cell_count, float min_occupancy, float min_velocity,
                                 float* out_records, int capacity);

#ifdef __cplusplus
}
#endif
#endif
