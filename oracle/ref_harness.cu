// ref_harness.cu — C entry points around the UNMODIFIED reference CUDA implementation (oracle/_ref).
// TEST INFRASTRUCTURE ONLY (tests/, bench.py --impl reference).  Built by oracle/build_ref.sh together with the
// reference's own sources where they lie under /root/reference (never copied into this repo):
//   dogm/src/dogm.cu, dogm/src/kernel/*.cu, dogm/demo/simulator/mapping/kernel/measurement_grid.cu
// against oracle/glm_shim (GLM is absent from this image).
//
// What the harness adds (its own code, not the reference's):
//   * ref_extract_noise: replays, on a COPY of the reference's XORWOW states and with the reference's launch
//     geometry, exactly the draws the next updateGrid will make (init_new_particles.cu:116-117, predict.cu:27-30,
//     init_new_particles.cu:178-187, resampling.cu:28) and returns them as per-slot buffers, so that another
//     implementation can be fed "the same pre-generated noise";
//   * stage-by-stage execution through the reference's public stage methods (dogm.h:148-156) with snapshots;
//   * ref_stage_resampling_verbose: dogm.cu:386-423 unrolled with the same thrust calls and reference kernels, so
//     that the joint CDF, the sorted draws and the ancestor indices (temporaries in the reference) become visible.
#define private public // updatePose / updateMeasurementGrid are private in dogm.h:141-146; the harness drives them singly
#include "dogm/dogm.h"
#undef private

#include "dogm/common.h"
#include "dogm/cuda_utils.h"
#include "dogm/dogm_types.h"
#include "dogm/kernel/resampling.h"
#include "mapping/kernel/measurement_grid.h"

#include <thrust/device_vector.h>
#include <thrust/sort.h>

#include <chrono>
#include <cstring>

struct ref_handle
{
    dogm::DOGM* dogm;
    curandState* scratch_states;
    int rng_threads;
    bool first_cycle_pending;
};

namespace
{

__global__ void extractInitKernel(curandState* __restrict__ global_state, float velocity, float2* out, int particle_count)
{ // draw order of initParticlesKernel2, init_new_particles.cu:101-124
    int thread_id = blockIdx.x * blockDim.x + threadIdx.x;
    int stride = blockDim.x * gridDim.x;
    curandState local_state = global_state[thread_id];
    for (int i = thread_id; i < particle_count; i += stride)
    {
        float vel_x = curand_uniform(&local_state, -velocity, velocity);
        float vel_y = curand_uniform(&local_state, -velocity, velocity);
        out[i] = make_float2(vel_x, vel_y);
    }
    global_state[thread_id] = local_state;
}

__global__ void extractPredictKernel(curandState* __restrict__ global_state, float process_noise_position,
                                     float process_noise_velocity, float4* out, int particle_count)
{ // draw order of predictKernel, predict.cu:16-52
    int thread_id = blockIdx.x * blockDim.x + threadIdx.x;
    int stride = blockDim.x * gridDim.x;
    curandState local_state = global_state[thread_id];
    for (int i = thread_id; i < particle_count; i += stride)
    {
        float noise_pos_x = curand_normal(&local_state, 0.0f, process_noise_position);
        float noise_pos_y = curand_normal(&local_state, 0.0f, process_noise_position);
        float noise_vel_x = curand_normal(&local_state, 0.0f, process_noise_velocity);
        float noise_vel_y = curand_normal(&local_state, 0.0f, process_noise_velocity);
        out[i] = make_float4(noise_pos_x, noise_pos_y, noise_vel_x, noise_vel_y);
    }
    global_state[thread_id] = local_state;
}

__global__ void extractBirthKernel(curandState* __restrict__ global_state, float stddev_velocity, float2* out,
                                   int particle_count)
{ // draw order of initNewParticlesKernel2, init_new_particles.cu:157-195
    int thread_id = blockIdx.x * blockDim.x + threadIdx.x;
    int stride = blockDim.x * gridDim.x;
    curandState local_state = global_state[thread_id];
    for (int i = thread_id; i < particle_count; i += stride)
    {
        float vel_x = curand_normal(&local_state, 0.0f, stddev_velocity);
        float vel_y = curand_normal(&local_state, 0.0f, stddev_velocity);
        out[i] = make_float2(vel_x, vel_y);
    }
    global_state[thread_id] = local_state;
}

__global__ void extractResampleKernel(curandState* __restrict__ global_state, float* out, int particle_count)
{ // draw order of resamplingGenerateRandomNumbersKernel, resampling.cu:17-32; stores (1 - u), the factor that
  // curand_uniform(state, 0, max) multiplies max with (cuda_utils.h:31-35)
    int thread_id = blockIdx.x * blockDim.x + threadIdx.x;
    int stride = blockDim.x * gridDim.x;
    curandState local_state = global_state[thread_id];
    for (int i = thread_id; i < particle_count; i += stride)
    {
        out[i] = 1.0f - curand_uniform(&local_state);
    }
    global_state[thread_id] = local_state;
}

void copy_block_to_host(void* dst, const dogm::ParticlesSoA& p)
{
    cudaMemcpy(dst, p.memory_block, (size_t)p.size * sizeof(dogm::Particle), cudaMemcpyDeviceToHost);
}

} // namespace

extern "C"
{

ref_handle* ref_create(const dogm::DOGM::Params* params)
{
    ref_handle* h = new ref_handle();
    h->dogm = new dogm::DOGM(*params);
    h->rng_threads = h->dogm->particles_grid.x * h->dogm->block_dim.x;
    cudaMalloc(&h->scratch_states, (size_t)h->rng_threads * sizeof(curandState));
    h->first_cycle_pending = true;
    // the reference leaves the particle sets uninitialised (dogm.cu:49-51); zero them so that runs are reproducible
    cudaMemset(h->dogm->particle_array.memory_block, 0, (size_t)h->dogm->particle_count * sizeof(dogm::Particle));
    cudaMemset(h->dogm->particle_array_next.memory_block, 0, (size_t)h->dogm->particle_count * sizeof(dogm::Particle));
    cudaMemset(h->dogm->birth_particle_array.memory_block, 0,
               (size_t)h->dogm->new_born_particle_count * sizeof(dogm::Particle));
    cudaMemset(h->dogm->grid_cell_array, 0, (size_t)h->dogm->grid_cell_count * sizeof(dogm::GridCell));
    cudaDeviceSynchronize();
    // redo initGridCellsKernel's effect after the memset (init.cu:69-84): indices -1
    std::vector<dogm::GridCell> cells((size_t)h->dogm->grid_cell_count);
    memset(cells.data(), 0, cells.size() * sizeof(dogm::GridCell));
    for (auto& c : cells)
    {
        c.start_idx = -1;
        c.end_idx = -1;
    }
    cudaMemcpy(h->dogm->grid_cell_array, cells.data(), cells.size() * sizeof(dogm::GridCell), cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    return h;
}

void ref_destroy(ref_handle* h)
{
    if (!h)
        return;
    cudaFree(h->scratch_states);
    delete h->dogm;
    delete h;
}

int ref_grid_size(ref_handle* h) { return h->dogm->grid_size; }
int ref_rng_thread_count(ref_handle* h) { return h->rng_threads; }
int ref_sizeof_particle(void) { return (int)sizeof(dogm::Particle); }
int ref_sizeof_grid_cell(void) { return (int)sizeof(dogm::GridCell); }

// the reference's own cycle, untouched (dogm.cu:115-131)
void ref_update_grid(ref_handle* h, void* meas, float x, float y, float yaw, float dt, int device)
{
    h->dogm->updateGrid((dogm::MeasurementCell*)meas, x, y, yaw, dt, device != 0);
    h->first_cycle_pending = false;
}

// wall-clock milliseconds of one updateGrid, the reference's own metric (demo/main.cpp:91-92, timer.cpp:12-15)
double ref_update_grid_timed(ref_handle* h, void* meas, float x, float y, float yaw, float dt, int device)
{
    auto t0 = std::chrono::steady_clock::now();
    h->dogm->updateGrid((dogm::MeasurementCell*)meas, x, y, yaw, dt, device != 0);
    auto t1 = std::chrono::steady_clock::now();
    h->first_cycle_pending = false;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// noise of the NEXT cycle (host buffers; init_vel may be NULL when the first cycle has already run)
void ref_extract_noise(ref_handle* h, int first_cycle, float* init_vel, float* predict, float* birth, float* resample_unit)
{
    dogm::DOGM& d = *h->dogm;
    const int N = d.particle_count, B = d.new_born_particle_count;
    cudaDeviceSynchronize();
    cudaMemcpy(h->scratch_states, d.rng_states, (size_t)h->rng_threads * sizeof(curandState), cudaMemcpyDeviceToDevice);
    float2 *d_init = nullptr, *d_birth = nullptr;
    float4* d_pred = nullptr;
    float* d_res = nullptr;
    cudaMalloc(&d_init, (size_t)(N > 0 ? N : 1) * sizeof(float2));
    cudaMalloc(&d_pred, (size_t)(N > 0 ? N : 1) * sizeof(float4));
    cudaMalloc(&d_birth, (size_t)(B > 0 ? B : 1) * sizeof(float2));
    cudaMalloc(&d_res, (size_t)(N > 0 ? N : 1) * sizeof(float));
    if (first_cycle)
        extractInitKernel<<<d.particles_grid, d.block_dim>>>(h->scratch_states, d.params.init_max_velocity, d_init, N);
    extractPredictKernel<<<d.particles_grid, d.block_dim>>>(h->scratch_states, d.params.stddev_process_noise_position,
                                                           d.params.stddev_process_noise_velocity, d_pred, N);
    extractBirthKernel<<<d.birth_particles_grid, d.block_dim>>>(h->scratch_states, d.params.stddev_velocity, d_birth, B);
    extractResampleKernel<<<d.particles_grid, d.block_dim>>>(h->scratch_states, d_res, N);
    cudaDeviceSynchronize();
    if (first_cycle && init_vel)
        cudaMemcpy(init_vel, d_init, (size_t)N * sizeof(float2), cudaMemcpyDeviceToHost);
    if (predict)
        cudaMemcpy(predict, d_pred, (size_t)N * sizeof(float4), cudaMemcpyDeviceToHost);
    if (birth)
        cudaMemcpy(birth, d_birth, (size_t)B * sizeof(float2), cudaMemcpyDeviceToHost);
    if (resample_unit)
        cudaMemcpy(resample_unit, d_res, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_init);
    cudaFree(d_pred);
    cudaFree(d_birth);
    cudaFree(d_res);
}

// ---- single stages ------------------------------------------------------------------------------------------
void ref_stage_update_measurement_grid(ref_handle* h, void* meas, int device)
{
    h->dogm->updateMeasurementGrid((dogm::MeasurementCell*)meas, device != 0);
    cudaDeviceSynchronize();
    h->first_cycle_pending = false;
}
void ref_stage_update_pose(ref_handle* h, float x, float y, float yaw)
{
    h->dogm->updatePose(x, y, yaw);
    cudaDeviceSynchronize();
}
void ref_stage_particle_prediction(ref_handle* h, float dt)
{
    h->dogm->particlePrediction(dt);
    cudaDeviceSynchronize();
}
void ref_stage_particle_assignment(ref_handle* h)
{
    h->dogm->particleAssignment();
    cudaDeviceSynchronize();
}
void ref_stage_grid_cell_occupancy_update(ref_handle* h, float dt)
{
    h->dogm->gridCellOccupancyUpdate(dt);
    cudaDeviceSynchronize();
}
void ref_stage_update_persistent_particles(ref_handle* h)
{
    h->dogm->updatePersistentParticles();
    cudaDeviceSynchronize();
}
void ref_stage_initialize_new_particles(ref_handle* h)
{
    h->dogm->initializeNewParticles();
    cudaDeviceSynchronize();
}
void ref_stage_statistical_moments(ref_handle* h)
{
    h->dogm->statisticalMoments();
    cudaDeviceSynchronize();
}

// dogm.cu:386-423 with the temporaries copied out.  cdf_out: N+B floats AFTER calc_resampled_indices (i.e. with the
// last entry replaced by the largest draw, resampling.cu:37-42); rand_sorted_out: N; idx_out: N.
void ref_stage_resampling_verbose(ref_handle* h, float* cdf_out, float* rand_sorted_out, int* idx_out, float* joint_max_out)
{
    dogm::DOGM& d = *h->dogm;
    const int particle_count = d.particle_count, new_born_particle_count = d.new_born_particle_count;

    thrust::device_ptr<float> persistent_weights(d.weight_array);
    thrust::device_ptr<float> new_born_weights(d.birth_particle_array.weight);

    thrust::device_vector<float> joint_weight_array;
    joint_weight_array.insert(joint_weight_array.end(), persistent_weights, persistent_weights + particle_count);
    joint_weight_array.insert(joint_weight_array.end(), new_born_weights, new_born_weights + new_born_particle_count);

    thrust::device_vector<float> joint_weight_accum(joint_weight_array.size());
    accumulate(joint_weight_array, joint_weight_accum);

    float joint_max = joint_weight_accum.back();

    dogm::resamplingGenerateRandomNumbersKernel<<<d.particles_grid, d.block_dim>>>(d.rand_array, d.rng_states, joint_max,
                                                                                 particle_count);
    thrust::device_ptr<float> rand_ptr(d.rand_array);
    thrust::device_vector<float> rand_vector(rand_ptr, rand_ptr + particle_count);
    thrust::sort(rand_vector.begin(), rand_vector.end());

    thrust::device_vector<int> idx_resampled(particle_count);
    dogm::calc_resampled_indices(joint_weight_accum, rand_vector, idx_resampled, joint_max);
    int* idx_array_resampled = thrust::raw_pointer_cast(idx_resampled.data());

    float new_weight = joint_max / particle_count;
    dogm::resamplingKernel<<<d.particles_grid, d.block_dim>>>(d.particle_array, d.particle_array_next,
                                                            d.birth_particle_array, idx_array_resampled, new_weight,
                                                            particle_count);
    cudaDeviceSynchronize();
    if (cdf_out)
        cudaMemcpy(cdf_out, thrust::raw_pointer_cast(joint_weight_accum.data()),
                   joint_weight_accum.size() * sizeof(float), cudaMemcpyDeviceToHost);
    if (rand_sorted_out)
        cudaMemcpy(rand_sorted_out, thrust::raw_pointer_cast(rand_vector.data()), (size_t)particle_count * sizeof(float),
                   cudaMemcpyDeviceToHost);
    if (idx_out)
        cudaMemcpy(idx_out, idx_array_resampled, (size_t)particle_count * sizeof(int), cudaMemcpyDeviceToHost);
    if (joint_max_out)
        *joint_max_out = joint_max;
}

void ref_stage_publish(ref_handle* h)
{ // dogm.cu:128-130
    h->dogm->particle_array = h->dogm->particle_array_next;
    cudaDeviceSynchronize();
}

// ---- state access ---------------------------------------------------------------------------------------------
void ref_get_particles(ref_handle* h, void* block) { copy_block_to_host(block, h->dogm->particle_array); }
void ref_get_particles_next(ref_handle* h, void* block) { copy_block_to_host(block, h->dogm->particle_array_next); }
void ref_get_birth_particles(ref_handle* h, void* block) { copy_block_to_host(block, h->dogm->birth_particle_array); }
void ref_set_particles(ref_handle* h, const void* block)
{
    cudaMemcpy(h->dogm->particle_array.memory_block, block, (size_t)h->dogm->particle_count * sizeof(dogm::Particle),
               cudaMemcpyHostToDevice);
}
void ref_get_grid_cells(ref_handle* h, void* out)
{
    cudaMemcpy(out, h->dogm->grid_cell_array, (size_t)h->dogm->grid_cell_count * sizeof(dogm::GridCell),
               cudaMemcpyDeviceToHost);
}
void ref_set_grid_cells(ref_handle* h, const void* in)
{
    cudaMemcpy(h->dogm->grid_cell_array, in, (size_t)h->dogm->grid_cell_count * sizeof(dogm::GridCell),
               cudaMemcpyHostToDevice);
}
void ref_get_meas_cells(ref_handle* h, void* out)
{
    cudaMemcpy(out, h->dogm->meas_cell_array, (size_t)h->dogm->grid_cell_count * sizeof(dogm::MeasurementCell),
               cudaMemcpyDeviceToHost);
}
void ref_get_weight_array(ref_handle* h, float* out)
{
    cudaMemcpy(out, h->dogm->weight_array, (size_t)h->dogm->particle_count * sizeof(float), cudaMemcpyDeviceToHost);
}
void ref_get_born_masses(ref_handle* h, float* out)
{
    cudaMemcpy(out, h->dogm->born_masses_array, (size_t)h->dogm->grid_cell_count * sizeof(float), cudaMemcpyDeviceToHost);
}
void ref_get_pose(ref_handle* h, float* out3)
{
    out3[0] = h->dogm->getPositionX();
    out3[1] = h->dogm->getPositionY();
    out3[2] = h->dogm->getYaw();
}
void* ref_device_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaMalloc(&p, bytes);
    return p;
}
void ref_device_free(void* p) { cudaFree(p); }
void ref_memcpy_h2d(void* dst, const void* src, size_t bytes) { cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice); }
int ref_last_error(void) { return (int)cudaGetLastError(); }

// ---- measurement grid, stage M1: the reference's createPolarGridTextureKernel on a plain CUDA surface ------------
int ref_polar_grid(const float* beams, int width, int height, float resolution, float stddev_range, float* out)
{
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float2>();
    cudaArray_t arr = nullptr;
    if (cudaMallocArray(&arr, &desc, width, height, cudaArraySurfaceLoadStore) != cudaSuccess)
        return -1;
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaSurfaceObject_t surf = 0;
    if (cudaCreateSurfaceObject(&surf, &res) != cudaSuccess)
        return -2;
    float* d_beams = nullptr;
    cudaMalloc(&d_beams, (size_t)width * sizeof(float));
    cudaMemcpy(d_beams, beams, (size_t)width * sizeof(float), cudaMemcpyHostToDevice);
    dim3 dim_block(32, 32);
    dim3 grid_dim(divUp(width, dim_block.x), divUp(height, dim_block.y));
    createPolarGridTextureKernel<<<grid_dim, dim_block>>>(surf, d_beams, width, height, resolution, stddev_range);
    cudaDeviceSynchronize();
    cudaMemcpy2DFromArray(out, (size_t)width * sizeof(float2), arr, 0, 0, (size_t)width * sizeof(float2), height,
                          cudaMemcpyDeviceToHost);
    cudaDestroySurfaceObject(surf);
    cudaFreeArray(arr);
    cudaFree(d_beams);
    return (int)cudaGetLastError();
}

// ---- multi-scan fusion: createPolarGridTextureKernel for the first scan, fusePolarGridTextureKernel for the others ----
int ref_polar_fused(const float* scans, int num_scans, int width, int height, float resolution, float stddev_range, float* out)
{
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float2>();
    cudaArray_t arr = nullptr;
    if (cudaMallocArray(&arr, &desc, width, height, cudaArraySurfaceLoadStore) != cudaSuccess)
        return -1;
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaSurfaceObject_t surf = 0;
    if (cudaCreateSurfaceObject(&surf, &res) != cudaSuccess)
        return -2;
    float* d_beams = nullptr;
    cudaMalloc(&d_beams, (size_t)width * sizeof(float));
    dim3 dim_block(32, 32);
    dim3 grid_dim(divUp(width, dim_block.x), divUp(height, dim_block.y));
    for (int s = 0; s < num_scans; s++)
    {
        cudaMemcpy(d_beams, scans + (size_t)s * width, (size_t)width * sizeof(float), cudaMemcpyHostToDevice);
        if (s == 0)
            createPolarGridTextureKernel<<<grid_dim, dim_block>>>(surf, d_beams, width, height, resolution, stddev_range);
        else
            fusePolarGridTextureKernel<<<grid_dim, dim_block>>>(surf, d_beams, width, height, resolution, stddev_range);
        cudaDeviceSynchronize();
    }
    cudaMemcpy2DFromArray(out, (size_t)width * sizeof(float2), arr, 0, 0, (size_t)width * sizeof(float2), height,
                          cudaMemcpyDeviceToHost);
    cudaDestroySurfaceObject(surf);
    cudaFreeArray(arr);
    cudaFree(d_beams);
    return (int)cudaGetLastError();
}

} // extern "C"
