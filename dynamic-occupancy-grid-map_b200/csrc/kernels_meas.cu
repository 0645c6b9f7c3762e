// Measurement grid from one lidar scan, as a single CUDA kernel.
//
// Replaces the reference's three-step path (demo/simulator/mapping/laser_to_meas_grid.cu:25-70):
//   createPolarGridTextureKernel  (kernel/measurement_grid.cu:33-89)   inverse sensor model -> polar RG32F GL texture
//   OpenGL triangle-fan render    (opengl/renderer.cpp:11-31,50-54,66-84; opengl/shader.cpp:22-35)  polar -> cartesian
//   cartesianGridToMeasurementGridKernel (kernel/measurement_grid.cu:115-132)  RGBA32F framebuffer -> MeasurementCell[]
// Every cartesian cell finds its fan triangle, interpolates the texture coordinate exactly as the rasteriser would,
// and takes a bilinear sample (mip level 0) of the inverse sensor model; no GL.
// The geometry of that lookup (fan triangle, texel pair, bilinear weights) depends only on the sensor, the grid and the
// number of beams, so it is computed once per beam count (k_meas_geom) and a scan costs two small kernels:
// k_meas_polar (inverse sensor model, K x H texels, L2-resident) and k_meas_apply (16 B in, 4 cached texels, 16 B out).
#include "dogm_internal.cuh"

#include <cmath>
#include <vector>

struct dogm_meas_handle
{
    dogm_laser_params params;
    float grid_length;
    float resolution;
    int gs;
    int H; // polar range bins
    int a0, wedges;
    float radius, start_angle;
    int max_beams;
    float* d_beams;
    float2* d_verts; // arc vertices (NDC offsets from the fan centre), wedges + 1 entries
    dogm_meas_cell* d_grid;
    float* h_beams_pinned;      // two staging halves of max_beams floats each, used in turn
    cudaEvent_t beams_copied[2]; // recorded behind the host-to-device copy out of each half
    bool beams_event_ready;
    int beams_half;
    float4* d_geom;  // per cartesian cell: (i0, j0, wu, wv) of its bilinear lookup, i0 = kNoTexel outside the fan
    int geom_beams;  // beam count d_geom was built for (0 = none)
    float2* d_polar; // K x H polar table of the current scan
    size_t polar_capacity;
    cudaStream_t stream;
};

namespace dogm_b200
{

struct MeasArgs
{
    int gs, K, H;
    int a0, wedges;
    float radius, start_angle, fov;
    float resolution, stddev_range;
    const float* beams;
    const float2* verts;
    dogm_meas_cell* out;
};

// inverse_sensor_model + clamp, measurement_grid.cu:33-70,80-85
__device__ __forceinline__ float2 polar_cell(const MeasArgs& a, float zk, int i)
{
    const float free_p = 0.15f + (float)i * (1.0f - 0.15f) / (float)a.H;
    float occ_m, free_m;
    if (isfinite(zk))
    {
        const int r = __float2int_rz(zk / a.resolution);
        const float diff = (float)(i - r) * a.resolution;
        const float occ = 0.95f * expf(-0.5f * diff * diff / (a.stddev_range * a.stddev_range));
        if (i <= r)
        {
            const bool o = occ > free_p;
            occ_m = o ? occ : 0.0f;
            free_m = o ? 0.0f : 1.0f - free_p;
        }
        else
        {
            occ_m = occ > 0.5f ? occ : 0.0f;
            free_m = 0.0f;
        }
    }
    else
    {
        occ_m = 0.0f;
        free_m = 1.0f - free_p;
    }
    const float eps = 0.00001f;
    return make_float2(fmaxf(eps, fminf(1.0f - eps, occ_m)), fmaxf(eps, fminf(1.0f - eps, free_m)));
}

// Scan-independent part of the lookup: which fan triangle the cell centre falls into and where the rasteriser's
// interpolated texture coordinate lands in the K x H polar texture.
__global__ void __launch_bounds__(kBlock) k_meas_geom(MeasArgs a, float4* geom)
{
    pdl_prologue(K_MISC * 2);
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= a.gs * a.gs)
        return;
    const int px = c % a.gs;
    const int py = a.gs - 1 - c / a.gs; // framebuffer row of this grid row (flip of measurement_grid.cu:120)
    float4 g = make_float4(__int_as_float(kNoTexel), __int_as_float(kNoTexel), 0.0f, 0.0f);

    const float X = ((float)px + 0.5f) * 2.0f / (float)a.gs - 1.0f;
    const float Y = ((float)py + 0.5f) * 2.0f / (float)a.gs - 1.0f;
    const float dx = X, dy = Y + 1.0f; // fan centre at NDC (0,-1), renderer.cpp:54
    const float deg = atan2f(dy, dx) * (180.0f / 3.14159265358979323846f);
    int k = (int)floorf(deg - (float)a.a0);
    k = max(0, min(a.wedges - 1, k));
    float beta = 0.0f, gamma = 0.0f;
    bool found = false;
    for (int attempt = 0; attempt < 3 && !found; attempt++)
    {
        const float2 e0 = a.verts[k], e1 = a.verts[k + 1];
        const float det = e0.x * e1.y - e0.y * e1.x;
        beta = (dx * e1.y - dy * e1.x) / det;
        gamma = (e0.x * dy - e0.y * dx) / det;
        if (gamma < 0.0f && k > 0)
            k -= 1;
        else if (beta < 0.0f && k < a.wedges - 1)
            k += 1;
        else
            found = true;
    }
    if (found && beta >= 0.0f && gamma >= 0.0f && beta + gamma <= 1.0f)
    {
        // texcoords: ((angle - start) / fov, 1) on the arc, (0,0) at the centre (renderer.cpp:13,29)
        const float s0 = ((float)(a.a0 + k) - a.start_angle) / a.fov;
        const float s1 = ((float)(a.a0 + k + 1) - a.start_angle) / a.fov;
        const float s = beta * s0 + gamma * s1;
        const float t = beta + gamma;
        const float u = 1.0f - s / (t + 1e-10f); // shader.cpp:29
        const float v = t;
        const float fu = u * (float)a.K - 0.5f;
        const float fv = v * (float)a.H - 0.5f;
        const float fu0 = floorf(fu), fv0 = floorf(fv);
        g = make_float4(__int_as_float((int)fu0), __int_as_float((int)fv0), fu - fu0, fv - fv0);
    }
    geom[c] = g;
}

// Per scan: bilinear sample (GL_LINEAR, mip level 0) of the polar table at the precomputed position.
__global__ void __launch_bounds__(kBlock)
    k_meas_apply(const float4* __restrict__ geom, const float2* __restrict__ table, int K, int H, int C, dogm_meas_cell* out)
{
    pdl_prologue(K_MEAS_GRID * 2);
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= C)
        return;
    const float4 g = __ldcs(geom + c);
    __stcs(reinterpret_cast<float4*>(out + c), meas_cell_from_polar(g, table, K, H));
}

__global__ void __launch_bounds__(kBlock) k_meas_polar(MeasArgs a, float2* out)
{
    pdl_prologue(K_MEAS_POLAR * 2);
    const int t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= a.K * a.H)
        return;
    const int b = t % a.K, i = t / a.K;
    out[t] = polar_cell(a, a.beams[b], i);
}

// Same table, the scan passed by value in the kernel parameters: no host-to-device copy node in front of the launch, so the
// programmatic launch chain of the stream runs on from the previous cycle (dogm_meas_generate_into, up to kBeamsByValue beams)
constexpr int kBeamsByValue = 960;
struct BeamBlock
{
    float z[kBeamsByValue];
};
// Nothing this kernel reads comes from the kernels in front of it in the stream (the scan is in its parameters; the last
// reader of the table, the cell kernel of the cycle before, finished several kernels ago), so it does its work first and waits
// for its predecessor only before it exits: its CTAs run in the tail of the previous cycle's resampling kernel, and completion
// still is transitive along the stream (this kernel ends after its predecessor has).
// (kEarly = false: the plain order - for a scan that replaces one no cycle has consumed yet (two early kernels could overlap), and
// whenever the host has not seen the last cell kernel complete: on a small grid a whole cycle's CTAs can be resident at once.)
template <bool kEarly>
__global__ void __launch_bounds__(kBlock) k_meas_polar_byvalue(MeasArgs a, float2* out, const __grid_constant__ BeamBlock scan)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!kEarly)
        asm volatile("griddepcontrol.wait;" ::: "memory");
    const int t = blockIdx.x * kBlock + threadIdx.x;
    if (t < a.K * a.H)
    {
        const int b = t % a.K, i = t / a.K;
        out[t] = polar_cell(a, scan.z[b], i);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && c_trace)
    {
        unsigned long long ts;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
        atomicMin(c_trace + K_MEAS_POLAR * 2, ts);
    }
}

// one more scan into the polar table: Dempster-Shafer combination of the table (prior) with the clamped inverse sensor model
// of the scan; the result is not clamped again (fusePolarGridTextureKernel + combine_masses, measurement_grid.cu:13-31,91-113)
__global__ void __launch_bounds__(kBlock) k_meas_polar_fuse(MeasArgs a, float2* table)
{
    pdl_prologue(K_MEAS_POLAR * 2 + 1);
    const int t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= a.K * a.H)
        return;
    const int b = t % a.K, i = t / a.K;
    const float2 meas = polar_cell(a, a.beams[b], i);
    const float2 prior = table[t];
    const float occ = prior.x, fre = prior.y;
    const float unknown_pred = 1.0f - occ - fre;
    const float meas_unknown = 1.0f - meas.x - meas.y;
    const float K = fre * meas.x + occ * meas.y;
    table[t] = make_float2((occ * meas_unknown + unknown_pred * meas.x + occ * meas.x) / (1.0f - K),
                           (fre * meas_unknown + unknown_pred * meas.y + fre * meas.y) / (1.0f - K));
}

static MeasArgs make_args(const dogm_meas_handle* m, int K, dogm_meas_cell* out)
{
    MeasArgs a;
    a.gs = m->gs;
    a.K = K;
    a.H = m->H;
    a.a0 = m->a0;
    a.wedges = m->wedges;
    a.radius = m->radius;
    a.start_angle = m->start_angle;
    a.fov = m->params.fov;
    a.resolution = m->params.resolution;
    a.stddev_range = m->params.stddev_range;
    a.beams = m->d_beams;
    a.verts = m->d_verts;
    a.out = out;
    return a;
}

static int upload_beams(dogm_meas_handle* m, const float* beams_host, int K, cudaStream_t stream)
{
    if (K <= 0 || !beams_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (!m->beams_event_ready)
    {
        DOGM_CHECK(cudaEventCreateWithFlags(&m->beams_copied[0], cudaEventDisableTiming));
        DOGM_CHECK(cudaEventCreateWithFlags(&m->beams_copied[1], cudaEventDisableTiming));
        m->beams_event_ready = true;
    }
    if (K > m->max_beams)
    {
        cudaStreamSynchronize(stream);
        cudaEventSynchronize(m->beams_copied[0]);
        cudaEventSynchronize(m->beams_copied[1]);
        cudaFree(m->d_beams);
        cudaFreeHost(m->h_beams_pinned);
        m->max_beams = K;
        DOGM_CHECK(cudaMalloc(&m->d_beams, (size_t)K * sizeof(float)));
        DOGM_CHECK(cudaMallocHost(&m->h_beams_pinned, 2 * (size_t)K * sizeof(float)));
    }
    // The copy out of the staging memory is asynchronous: in a pipelined loop (generate_into, updateGridAsync, generate_into,
    // ...) it may still be queued behind the previous cycle when the next scan arrives.  Two halves used in turn, and the
    // host waits for the copy out of a half (issued two scans ago) before it writes that half again.
    const int half = m->beams_half;
    m->beams_half ^= 1;
    DOGM_CHECK(cudaEventSynchronize(m->beams_copied[half]));
    float* stage = m->h_beams_pinned + (size_t)half * m->max_beams;
    memcpy(stage, beams_host, (size_t)K * sizeof(float));
    DOGM_CHECK(cudaMemcpyAsync(m->d_beams, stage, (size_t)K * sizeof(float), cudaMemcpyHostToDevice, stream));
    DOGM_CHECK(cudaEventRecord(m->beams_copied[half], stream));
    return 0;
}

// grows the polar table / rebuilds the cached geometry when the beam count changes
static int prepare_scan(dogm_meas_handle* m, int K, cudaStream_t stream)
{
    const long long C = (long long)m->gs * m->gs;
    const size_t texels = (size_t)K * m->H;
    if (texels > m->polar_capacity)
    {
        cudaStreamSynchronize(stream);
        cudaFree(m->d_polar);
        m->d_polar = nullptr;
        m->polar_capacity = 0;
        DOGM_CHECK(cudaMalloc(&m->d_polar, texels * sizeof(float2)));
        m->polar_capacity = texels;
    }
    if (m->geom_beams != K)
    {
        const MeasArgs a = make_args(m, K, nullptr);
        launch_chained(stream, k_meas_geom, div_up(C, kBlock), kBlock, 0, a, m->d_geom);
        DOGM_CHECK(cudaGetLastError());
        m->geom_beams = K;
    }
    return 0;
}

// polar table of the scan in d_beams (first = overwrite, otherwise fuse into the table)
static int launch_polar(dogm_meas_handle* m, int K, bool first, cudaStream_t stream, dogm_handle* timing)
{
    const size_t texels = (size_t)K * m->H;
    const MeasArgs a = make_args(m, K, nullptr);
    if (timing)
    {
        LaunchScope ls(timing, K_MEAS_POLAR, 8.0 * (double)texels);
        if (first)
            launch_chained(stream, k_meas_polar, div_up((long long)texels, kBlock), kBlock, 0, a, m->d_polar);
        else
            launch_chained(stream, k_meas_polar_fuse, div_up((long long)texels, kBlock), kBlock, 0, a, m->d_polar);
    }
    else if (first)
        launch_chained(stream, k_meas_polar, div_up((long long)texels, kBlock), kBlock, 0, a, m->d_polar);
    else
        launch_chained(stream, k_meas_polar_fuse, div_up((long long)texels, kBlock), kBlock, 0, a, m->d_polar);
    return (int)cudaGetLastError();
}

// cartesian resampling of the polar table
static int launch_apply(dogm_meas_handle* m, int K, dogm_meas_cell* out, cudaStream_t stream, dogm_handle* timing)
{
    const long long C = (long long)m->gs * m->gs;
    if (timing)
    {
        LaunchScope ls(timing, K_MEAS_GRID, 32.0 * (double)C);
        launch_chained(stream, k_meas_apply, div_up(C, kBlock), kBlock, 0, m->d_geom, m->d_polar, K, m->H, (int)C, out);
    }
    else
        launch_chained(stream, k_meas_apply, div_up(C, kBlock), kBlock, 0, m->d_geom, m->d_polar, K, m->H, (int)C, out);
    return (int)cudaGetLastError();
}

static int launch_scan(dogm_meas_handle* m, int K, dogm_meas_cell* out, cudaStream_t stream, dogm_handle* timing)
{
    int e = prepare_scan(m, K, stream);
    e = e ? e : launch_polar(m, K, true, stream, timing);
    e = e ? e : launch_apply(m, K, out, stream, timing);
    return e;
}

int materialize_meas(dogm_handle* h)
{
    if (!h->lazy_meas.pending)
        return 0;
    h->lazy_meas.pending = false;
    // a reader of the polar table is in flight now: the next scan's table kernel must not run ahead of its dependency wait
    // until the host has seen the stream drain (the flag is only set again where the host observes a synchronisation)
    h->cell_kernel_done = false;
    LaunchScope ls(h, K_MEAS_GRID, 32.0 * (double)h->C);
    launch_chained(h->stream, k_meas_apply, div_up(h->C, kBlock), kBlock, 0, h->lazy_meas.geom, h->lazy_meas.polar, h->lazy_meas.K,
                   h->lazy_meas.H, h->C, h->meas);
    return (int)cudaGetLastError();
}

int trace_bind_meas(unsigned long long* p)
{
    return (int)cudaMemcpyToSymbol(c_trace, &p, sizeof(p));
}

} // namespace dogm_b200

using namespace dogm_b200;

extern "C" int dogm_meas_create(const dogm_laser_params* params, float grid_length, float resolution,
                                dogm_meas_handle** out)
{
    if (!params || !out)
        return DOGM_ERR_INVALID_ARGUMENT;
    dogm_meas_handle* m = new dogm_meas_handle();
    m->params = *params;
    m->grid_length = grid_length;
    m->resolution = resolution;
    m->gs = (int)(grid_length / resolution);                 // laser_to_meas_grid.cu:11
    m->H = (int)(params->max_range / params->resolution);    // laser_to_meas_grid.cu:35
    m->radius = 2.0f * (params->max_range / grid_length);    // renderer.cpp:52
    const float half_fov = params->fov / 2.0f;               // renderer.cpp:15-17
    m->start_angle = 90.0f - half_fov;
    const float end_angle = 90.0f + half_fov;
    m->a0 = (int)m->start_angle;
    std::vector<float2> verts;
    for (int angle = m->a0; (float)angle <= end_angle; angle++) // renderer.cpp:19-30
    {
        const float rad = (float)((double)angle * 3.14159265358979323846 / 180.0);
        verts.push_back(make_float2(m->radius * cosf(rad), m->radius * sinf(rad)));
    }
    m->wedges = (int)verts.size() - 1;
    m->max_beams = 0;
    m->d_beams = nullptr;
    m->h_beams_pinned = nullptr;
    m->beams_event_ready = false;
    m->beams_half = 0;
    m->d_verts = nullptr;
    m->d_grid = nullptr;
    m->d_geom = nullptr;
    m->geom_beams = 0;
    m->d_polar = nullptr;
    m->polar_capacity = 0;
    if (m->gs <= 0 || m->H <= 0 || m->wedges < 1)
    {
        delete m;
        return DOGM_ERR_INVALID_ARGUMENT;
    }
    DOGM_CHECK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    DOGM_CHECK(cudaMalloc(&m->d_verts, verts.size() * sizeof(float2)));
    DOGM_CHECK(cudaMemcpy(m->d_verts, verts.data(), verts.size() * sizeof(float2), cudaMemcpyHostToDevice));
    DOGM_CHECK(cudaMalloc(&m->d_grid, (size_t)m->gs * m->gs * sizeof(dogm_meas_cell)));
    DOGM_CHECK(cudaMalloc(&m->d_geom, (size_t)m->gs * m->gs * sizeof(float4)));
    *out = m;
    return 0;
}

extern "C" void dogm_meas_destroy(dogm_meas_handle* m)
{
    if (!m)
        return;
    cudaStreamSynchronize(m->stream);
    cudaFree(m->d_beams);
    cudaFreeHost(m->h_beams_pinned);
    if (m->beams_event_ready)
    {
        cudaEventDestroy(m->beams_copied[0]);
        cudaEventDestroy(m->beams_copied[1]);
    }
    cudaFree(m->d_verts);
    cudaFree(m->d_grid);
    cudaFree(m->d_geom);
    cudaFree(m->d_polar);
    cudaStreamDestroy(m->stream);
    delete m;
}

extern "C" int dogm_meas_get_grid_size(const dogm_meas_handle* m)
{
    return m ? m->gs : 0;
}

extern "C" int dogm_meas_generate(dogm_meas_handle* m, const float* beam_ranges_host, int num_beams,
                                  dogm_meas_cell** out_device)
{
    if (!m)
        return DOGM_ERR_INVALID_ARGUMENT;
    int e = upload_beams(m, beam_ranges_host, num_beams, m->stream);
    if (e)
        return e;
    e = launch_scan(m, num_beams, m->d_grid, m->stream, nullptr);
    if (e)
        return e;
    DOGM_CHECK(cudaStreamSynchronize(m->stream)); // laser_to_meas_grid.cu:67
    if (out_device)
        *out_device = m->d_grid;
    return 0;
}

extern "C" int dogm_meas_generate_into(dogm_meas_handle* m, dogm_handle* h, const float* beam_ranges_host, int num_beams)
{
    if (!m || !h || m->gs != h->gs || num_beams <= 0 || !beam_ranges_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (h->band.enabled)
        return DOGM_ERR_UNSUPPORTED; // a band holds some rows of the grid only: generate the grid, hand the band its rows
    const int K = num_beams;
    int e = 0;
    if (K <= kBeamsByValue)
    {
        if ((e = prepare_scan(m, K, h->stream)))
            return e;
        BeamBlock scan;
        memcpy(scan.z, beam_ranges_host, (size_t)K * sizeof(float));
        const MeasArgs a = make_args(m, K, nullptr);
        LaunchScope ls(h, K_MEAS_POLAR, 8.0 * (double)K * m->H);
        if (h->lazy_meas.pending || !h->cell_kernel_done)
            launch_chained(h->stream, k_meas_polar_byvalue<false>, div_up((long long)K * m->H, kBlock), kBlock, 0, a, m->d_polar, scan);
        else
            launch_chained(h->stream, k_meas_polar_byvalue<true>, div_up((long long)K * m->H, kBlock), kBlock, 0, a, m->d_polar, scan);
        e = (int)cudaGetLastError();
    }
    else
    {
        e = upload_beams(m, beam_ranges_host, K, h->stream);
        e = e ? e : prepare_scan(m, K, h->stream);
        e = e ? e : launch_polar(m, K, true, h->stream, h);
    }
    if (e)
        return e;
    // the cartesian resampling is left to the cell kernel of the next cycle (run_occupancy_update), or to materialize_meas()
    h->lazy_meas.pending = true;
    h->lazy_meas.geom = m->d_geom;
    h->lazy_meas.polar = m->d_polar;
    h->lazy_meas.K = K;
    h->lazy_meas.H = m->H;
    if (!h->first_measurement_received)
        return materialize_meas(h); // the first cycle distributes the particles over this grid before any cell kernel runs
    return 0;
}

extern "C" int dogm_meas_generate_fused(dogm_meas_handle* m, const float* scans_host, int num_scans, int num_beams,
                                        dogm_meas_cell** out_device, float* out_polar_host)
{
    if (!m || !scans_host || num_scans < 1 || num_beams < 1)
        return DOGM_ERR_INVALID_ARGUMENT;
    for (int s = 0; s < num_scans; s++)
    {
        int e = upload_beams(m, scans_host + (size_t)s * num_beams, num_beams, m->stream);
        e = e ? e : (s == 0 ? prepare_scan(m, num_beams, m->stream) : 0);
        e = e ? e : launch_polar(m, num_beams, s == 0, m->stream, nullptr);
        if (e)
            return e;
    }
    int e = launch_apply(m, num_beams, m->d_grid, m->stream, nullptr);
    if (e)
        return e;
    if (out_polar_host)
        DOGM_CHECK(cudaMemcpyAsync(out_polar_host, m->d_polar, (size_t)num_beams * m->H * sizeof(float2), cudaMemcpyDeviceToHost,
                                   m->stream));
    DOGM_CHECK(cudaStreamSynchronize(m->stream));
    if (out_device)
        *out_device = m->d_grid;
    return 0;
}

extern "C" int dogm_meas_polar_grid(dogm_meas_handle* m, const float* beam_ranges_host, int num_beams, float* out_host)
{
    if (!m || !out_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    int e = upload_beams(m, beam_ranges_host, num_beams, m->stream);
    if (e)
        return e;
    float2* d_out = nullptr;
    const size_t count = (size_t)num_beams * m->H;
    DOGM_CHECK(cudaMalloc(&d_out, count * sizeof(float2)));
    const MeasArgs a = make_args(m, num_beams, nullptr);
    launch_chained(m->stream, k_meas_polar, div_up((long long)count, kBlock), kBlock, 0, a, d_out);
    DOGM_CHECK(cudaGetLastError());
    DOGM_CHECK(cudaMemcpyAsync(out_host, d_out, count * sizeof(float2), cudaMemcpyDeviceToHost, m->stream));
    DOGM_CHECK(cudaStreamSynchronize(m->stream));
    cudaFree(d_out);
    return 0;
}
