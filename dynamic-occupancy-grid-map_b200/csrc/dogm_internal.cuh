// Internal declarations of libdogm_b200: handle layout, device-side helpers, kernel launch entry points.
#pragma once

#include "../../include/dogm_b200.h"
#include "philox.cuh"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

namespace dogm_b200
{

// ---------------------------------------------------------------------------------------------------------
// error handling: same print format as the reference's CHECK_ERROR (cuda_utils.h:13-24), but the code is kept
// ---------------------------------------------------------------------------------------------------------
inline int check_cuda(cudaError_t code, const char* file, int line)
{
    if (code != cudaSuccess)
        printf("GPU Kernel Error: %s %s %d\n", cudaGetErrorString(code), file, line);
    return (int)code;
}
#define DOGM_CHECK(expr)                                                                                               \
    do                                                                                                                 \
    {                                                                                                                  \
        int _e = ::dogm_b200::check_cuda((expr), __FILE__, __LINE__);                                                  \
        if (_e != 0)                                                                                                   \
            return _e;                                                                                                 \
    } while (0)

inline int div_up(long long a, long long b)
{
    return (int)((a + b - 1) / b);
}

// ---------------------------------------------------------------------------------------------------------
// geometry constants
// ---------------------------------------------------------------------------------------------------------
constexpr int kBlock = 256;           // threads per CTA of every streaming kernel
constexpr int kWarpsPerBlock = 8;
constexpr int kWideBlock = 1024;      // CTA size of the kernels that own a whole sort tile / a 32-bin column
constexpr int kTileItems = 4096;      // particles per sort tile (one CTA, 16 per thread, 512 per warp)
constexpr int kItemsPerThread = kTileItems / kBlock;
constexpr int kRoundsPerWarp = kTileItems / kWarpsPerBlock / 32;
constexpr int kMaxDigitBits = 12;     // <= 4096 bins: 8 warps * 4096 * 2 B = 64 KB of rank counters
constexpr int kMaxPasses = 3;
constexpr int kSegChunk = 256;        // particles per warp in the segmented reduction
constexpr int kCellBlock = 256;       // cells per CTA in the cell kernels (also the granularity of the born-mass scan)
constexpr int kCdfItems = 16;         // CDF entries per thread
constexpr int kCdfTile = kBlock * kCdfItems;

// ---------------------------------------------------------------------------------------------------------
// device-side views
// ---------------------------------------------------------------------------------------------------------
struct ParticleSet // ParticlesSoA, dogm_types.h:51-144: one block of n*28 bytes
{
    void* block;
    float4* state;
    int* idx;
    float* weight;
    uint8_t* assoc;
    int n;
};

inline void particle_set_assign(ParticleSet& p, void* block, int n)
{
    p.block = block;
    p.n = n;
    char* b = (char*)block;
    p.state = (float4*)(b + DOGM_PARTICLE_STATE_OFFSET(n));
    p.idx = (int*)(b + DOGM_PARTICLE_IDX_OFFSET(n));
    p.weight = (float*)(b + DOGM_PARTICLE_WEIGHT_OFFSET(n));
    p.assoc = (uint8_t*)(b + DOGM_PARTICLE_ASSOC_OFFSET(n));
}

struct __align__(16) PRec // a particle inside a cycle: one 32-byte DRAM sector, written once by the prediction
{
    float x, y;     // lo half: position, cell, flag
    int key;        // grid_cell_idx
    uint32_t assoc; // associated flag
    float vx, vy;   // hi half: what the per-cell sums need
    float weight;
    uint32_t pad;
};

struct __align__(16) CellSums // per non-empty cell, written by the segmented reduction
{
    float s0; // sum w           (accumulated in double, rounded once)
    float s1; // sum w*vx
    float s2; // sum w*vy
    float s3; // sum (w*vx)*vx
    float s4; // sum (w*vy)*vy
    float s5; // sum (w*vx)*vy
    float pad0, pad1;
};

struct __align__(16) SegPiece // open piece of a segment at a chunk border
{
    double s0;
    float s1, s2, s3, s4, s5;
    int pad;
};

enum : int
{
    SEG_LEAD = 1,    // first particle of the chunk continues a segment of the previous chunk
    SEG_TRAIL = 2,   // last particle of the chunk continues into the next chunk
    SEG_THROUGH = 4, // the whole chunk is the inside of one segment
};

struct DeviceScalars // device-resident scalars: nothing is read back by the host inside a cycle
{
    double born_total;   // sum of born masses (normaliser of the birth / init slot distribution)
    double weight_total; // last entry of the joint weight CDF
};

// ---------------------------------------------------------------------------------------------------------
// band mode, device-paced cycle: what the bands tell each other travels GPU to GPU (peer stores into the receiver's
// mailbox), what a band derives from it stays in its device-resident counts - the host enqueues a whole cycle and
// looks at the counts once, at the end
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxBands = 16;

enum : int
{
    BAND_ERR_SEND_OVERFLOW = 1,  // more particles crossed an edge than the exchange boxes hold
    BAND_ERR_PARTICLE_CAP = 2,   // the band's particles do not fit its capacity
    BAND_ERR_BIRTH_CAP = 4,      // its share of the birth particles does not fit
    BAND_ERR_TIMEOUT = 8,        // a neighbour's message did not arrive
};

struct BandCounts
{
    int n_sort;           // particles after the neighbours' records have joined (sort, per-cell sums, CDF)
    int n_lo, n_hi;       // records received from the lower / upper neighbour
    int sent_lo, sent_hi; // records sent
    int B;                // birth particles of this band
    int birth_slot_base;  // global number of its first birth slot
    int n_out;            // resampled particles of this band
    int err;              // BAND_ERR_* bits
    int pad;
    long long out_base;   // global number of its first output slot
    double born_local, born_base, born_total;
    double weight_local, cdf_base, weight_total;
};

struct BandMail // lives on the band's GPU, written by the other bands' kernels (one writer per word)
{
    unsigned long long sent_from_lo; // (sequence number << 32) | records in the lower neighbour's upper send box
    unsigned long long sent_from_hi;
    unsigned long long pad[2];
    double born[kMaxBands];          // normaliser shares of all bands, tagged with the cycle's sequence number
    double weight[kMaxBands];
    unsigned int born_seq[kMaxBands];
    unsigned int weight_seq[kMaxBands];
};

// first and one-past-last slot of a band in a slot numbering that spans all bands: slots are handed out in proportion
// to mass, slot end of a prefix = int(float(prefix) * scale) exactly as the kernels evaluate it per cell
__host__ __device__ inline void band_slot_range(double before, double local, double total, int count_glob, int* first, int* past)
{
    if (!(total > 0.0))
    {
        *first = *past = 0;
        return;
    }
    const float scale = (float)count_glob / (float)total;
    *first = (int)((float)before * scale);
    *past = (int)((float)(before + local) * scale);
    if (*past < *first)
        *past = *first;
}

// the output slots of the whole grid whose resampling offsets fall into a band's part of the global CDF
__host__ __device__ inline void band_output_range(uint64_t seed, uint32_t cycle, bool systematic, long long n_glob, double weight_before,
                                                  double weight_local, double weight_total, long long* first, long long* past)
{
    long long i_lo = 0, i_hi = 0;
    if (weight_total > 0.0 && n_glob > 0 && weight_local > 0.0)
    {
        const double step = weight_total / (double)n_glob;
        const uint32_t s_lo = (uint32_t)seed, s_hi = (uint32_t)(seed >> 32);
        const float u0 = systematic ? u01_half_open(philox4x32_10(0u, STAGE_RESAMPLE, cycle, 0u, s_lo, s_hi).x) : 0.0f;
        // smallest slot whose offset ((i + u) * step, the expression of resample_offset) exceeds x; offsets ascend with i
        const double bound[2] = {weight_before, weight_before + weight_local}; // (== the next band's weight_before: same addition)
        long long found[2];
        for (int k = 0; k < 2; k++)
        {
            long long lo = 0, hi = n_glob;
            while (lo < hi)
            {
                const long long mid = lo + ((hi - lo) >> 1);
                const float u = systematic ? u0 : u01_half_open(philox4x32_10((uint32_t)mid, STAGE_RESAMPLE, cycle, 0u, s_lo, s_hi).x);
                if (((double)mid + (double)u) * step > bound[k])
                    hi = mid;
                else
                    lo = mid + 1;
            }
            found[k] = lo;
        }
        i_lo = weight_before > 0.0 ? found[0] : 0;
        i_hi = bound[1] >= weight_total ? n_glob : found[1];
        if (i_hi < i_lo)
            i_hi = i_lo;
    }
    *first = i_lo;
    *past = i_hi;
}

struct CycleShift // ego-motion compensation of this cycle (dogm.cu:175-178)
{
    int active;
    int x_move;
    int y_move;
};

// ---------------------------------------------------------------------------------------------------------
// kernel timing (bench instrumentation)
// ---------------------------------------------------------------------------------------------------------
enum KernelId : int
{
    K_PREDICT = 0,
    K_TILE_HIST,
    K_HIST_SCAN,
    K_SCATTER,
    K_SEGSUM,
    K_SEGFIX,
    K_CELL,
    K_BLOCKSUM_SCAN,
    K_WEIGHTS,
    K_MEAS_POLAR,
    K_BIRTH_PARTICLES,
    K_CDF_CHAIN,
    K_UNUSED,
    K_RESAMPLE,
    K_INIT_MASSES,
    K_INIT_PARTICLES,
    K_MEAS_GRID,
    K_MISC,
    K_MEMSET,
    K_COUNT
};

static const char* const kKernelNames[K_COUNT] = {
    "k_predict",       "k_tile_hist",       "k_hist_scan",  "k_scatter",   "k_segsum",      "k_segfix",
    "k_cell",          "k_blocksum_scan",   "k_weights",    "k_meas_polar", "k_birth_particles", "k_cdf_chain",
    "k_unused",       "k_resample",        "k_init_masses", "k_init_particles", "k_meas_grid", "k_misc",
    "memset"};

struct TimedLaunch
{
    int id;
    cudaEvent_t e0, e1;
};

} // namespace dogm_b200

// ---------------------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------------------
struct dogm_handle
{
    dogm_params params;
    dogm_options opts;
    int gs, C, N, B;              // row length, cells, current persistent / birth particle counts
    // Band mode (dogm_create_band): this handle owns the rows [row0, row0 + rows) of a G x G grid and the particles that
    // lie there; particle counts vary from cycle to cycle within the capacities; an orchestrator exchanges the
    // particles that cross a band edge and the global normalisers.  Without bands: G = rows = gs, row0 = 0, the
    // capacities are the counts, the bases are 0 - every formula below then reduces to the single-grid one.
    struct Band
    {
        int enabled;
        int G, row0, rows;
        int n_cap, b_cap;          // buffer capacities (particles)
        int n_glob, b_glob;        // particle counts of the whole grid (Params)
        uint64_t salt;             // added to the seed for per-particle noise (ranks must not share their streams)
        // set per cycle by the orchestrator
        double born_base;          // born mass of the bands before this one
        int birth_slot_base;       // global number of this band's first birth slot
        double cdf_base;           // joint weight of the bands before this one
        long long out_base;        // global number of this band's first resampled particle
        int n_out;                 // resampled particles of this band
        // migration and halo buffers
        dogm_b200::PRec* send[2];  // particles leaving through the lower / upper edge (records with global coordinates)
        dogm_b200::PRec* recv[2];
        uint8_t* mig;              // per slot: leaves through the lower (1) / upper (2) edge (k_predict -> k_outbox_*)
        double* out_cnt[2];        // per tile of 4096 slots: particles leaving through either edge
        double* out_off[2];        // their exclusive prefix
        double* out_total;         // device, 2 totals
        double* pin;               // pinned host staging for the scalars a phase hands back (8 doubles)
        int send_cap;
        float* halo[2];            // rows of the neighbours' previous free masses needed by an ego-motion shift in y
        int halo_rows;
        int halo_valid;            // the orchestrator filled the halo rows for this cycle
        double born_local, weight_local; // host copies of this band's normaliser shares
        // device-paced cycle (dogm_band_group): see BandCounts / BandMail
        dogm_b200::BandCounts* cnt;       // device
        dogm_b200::BandCounts* cnt_host;  // pinned copy, valid after the end-of-cycle synchronisation
        dogm_b200::BandMail* mail;        // this band's mailbox (device)
        dogm_b200::BandMail* peer_mail[dogm_b200::kMaxBands]; // all bands' mailboxes in band order (peer-accessible)
        int group_size, group_rank;
        const dogm_b200::BandCounts* dev_cnt; // non-null while the stages of a device-paced cycle are being enqueued
        uint32_t seq;                      // sequence number of the running cycle's messages
        int n_pred;                        // particle count the running cycle's prediction worked on
        // launch sizes of the device-paced cycle: estimates from the previous cycle (the kernels loop, so any value is safe)
        int est_recv, est_birth, est_out, est_forced;
        bool profile;                      // record events at the stage boundaries (dogm_band_set_profile)
        cudaEvent_t stage_ev[8];
        float stage_ms[7];                 // work and wait intervals of the last profiled cycle (dogm_band_stage_times)
    } band;
    int device;
    int sm_count;
    cudaStream_t stream;

    // reference-visible buffers (dogm.h:159-191)
    dogm_b200::ParticleSet pa;    // particle_array (== particle_array_next: the next population is written in place)
    dogm_b200::PRec* rec;         // the particles of the running cycle as 32-byte records in slot order
    int* key0;                    // their cell indices in slot order (input of the first sort pass)
    int2* pairs[2];               // (cell, slot) pairs: ping-pong buffers of the sort passes
    int2* spair;                  // the sorted pairs (one of pairs[]): position p of the sorted set is record spair[p].y
    float* sw;                    // predicted weights in sorted order (compact copy, written by the segmented reduction)
    bool pa_current;              // the SoA block holds the current particles (false between prediction and resampling)
    bool rec_valid;               // rec / key0 hold the current particles
    bool weights_deferred;        // weight_array of this cycle is still to be produced (by the fused CDF kernels)
    bool sorted_valid;            // spair / sw describe the current particles sorted by cell and the per-cell sums exist
    dogm_b200::ParticleSet birth; // birth_particle_array
    dogm_grid_cell* grid;
    // pipelined read-out (dogm_get_grid_cells_begin / _wait): a second GridCell buffer the cell kernel alternates with, so that
    // the copy of cycle n's cells to the host (copy stream) runs under cycle n + 1; null until the first _begin
    dogm_grid_cell* grid_alt;
    cudaStream_t copy_stream;
    cudaEvent_t cycle_done_ev;     // main stream: everything enqueued before the read-out was started
    cudaEvent_t grid_copy_ev[2];   // copy stream: the read-out of `grid` (0) / of the buffer now in `grid_alt` (1) has finished
    bool grid_copy_busy[2];
    bool grid_swap_pending;        // the next cell kernel writes the other buffer (set by _begin, done by run_occupancy_update)
    dogm_meas_cell* meas;
    float* weight_array;
    float* born_masses;
    const dogm_meas_cell* meas_src; // caller's device measurement grid of the running cycle (copied by the cell kernel)
    // Measurement grid still to be produced (dogm_meas_generate_into): the polar table of the scan exists, the cartesian
    // resampling (k_meas_apply) is left to the cell kernel of the next cycle, which writes `meas` on the way - or to
    // materialize_meas() when somebody wants to see `meas` earlier.  geom / polar belong to the generator handle.
    struct
    {
        bool pending;
        const float4* geom;
        const float2* polar;
        int K, H;
    } lazy_meas;
    // the host has seen the last launched cell kernel complete (a synchronisation, or the published dynamic-cell list): only then
    // may the next scan's polar-table kernel run ahead of its dependency wait - on a small grid the CTAs of a whole cycle can be
    // resident at once, each waiting for its predecessor, and an early writer of the table could overtake the cell kernel reading it
    bool cell_kernel_done;
    bool ranges_in_soa;             // cell_start / cell_end hold the ranges of the last assignment (not yet consumed)

    // per-cell working set
    float* free_cur;  // GridCell.free_mass of the previous cycle (SoA copy; the cell kernel reads it at the shifted index)
    float* free_next; // written by the cell kernel; swapped with free_cur
    int* cell_start;
    int* cell_end;
    dogm_b200::CellSums* cell_sums;
    float4* cell_coef; // {likelihood, p_A*mu_A, (1-p_A)*mu_UA, over-unit divisor or 0} for non-empty cells
    double* cell_prefix; // block-local inclusive prefix (double) of the born / initial masses, per cell
    double* blk_sum;   // born-mass sum per 256-cell block
    double* blk_off;   // exclusive prefix of blk_sum: inside the block's group of 256 (birth kernel) or complete (k_blocksum_scan)
    int n_cell_blocks;
    // Shortcut of the cell kernel for blocks of 256 cells in which nothing happens (no particles, no measurement, no free mass
    // left from the cycle before - most of a large grid): blk_age counts for how many cycles in a row a block's outputs have
    // been the empty record (0, 1, 2 = "for at least two": both free-mass buffers, the grid cells, born masses and prefix of
    // the block hold exactly what the kernel would write again).
    uint8_t* blk_age[2]; // read (previous cycle) / written (this cycle), swapped per cell-kernel launch
    bool quiet_off;       // the cell kernel runs without the shortcut (small grids; DOGM_B200_NO_QUIET)
    int n_blk_groups;  // groups of 256 blocks
    double* grp_word;  // [n_blk_groups] group sums published by the scanning CTAs of the birth kernel (parity words)
    double* grp_zero;  // [n_blk_groups + 1] zeros: the group offsets when blk_off holds complete offsets
    uint32_t birth_epoch;

    // counting sort
    int key_bits, passes;
    int digit_shift[dogm_b200::kMaxPasses];
    int digit_bins[dogm_b200::kMaxPasses];
    uint32_t* hist[dogm_b200::kMaxPasses];     // [tiles][bins] per pass
    uint32_t* bin_base[dogm_b200::kMaxPasses]; // chain words of k_hist_scan (bins / 32 x u64)
    uint32_t scan_epoch[dogm_b200::kMaxPasses];
    int tiles;
    bool hist0_valid; // pass-0 tile histograms were produced by the prediction kernel for the current keys
    bool hist0_bucket; // ... and they are histograms over the buckets of the bucket sort (not over the first radix digit)
    bool sort_fuse_off; // DOGM_B200_SORT=radix: keep the separate table scans / pair histogram for few tiles as well (A/B, tests)
    // Bucket sort (kernels_particles.cu, "Bucket sort"): the population enters a cycle in cell order, so the cells of every
    // 4096th particle are splitters for the new keys; grouping pass by bucket + one counting sort per bucket replace the radix
    // passes.  Used when the keys come from the prediction kernel of a handle without bands and the tables stay small.
    struct
    {
        bool enabled;
        bool samples_valid; // smp_raw was written by the last resampling (else k_sample_keys reads the particle block)
        int bins;           // buckets the tables are sized for (multiple of 32)
        int n_spl, stride_shift; // one splitter sample per (1 << stride_shift) slots
        int* smp_raw;       // [n_spl]
        uint16_t* bkt;      // [N] bucket number per slot
        int* org;           // [bins] first cell of every bucket
        int *plan_key, *plan_slot, *plan_base, *plan_pos; // [n_spl] the cycle's splitters (k_bucket_plan)
    } bucket;

    // segmented reduction
    int n_chunks;
    dogm_b200::SegPiece* seg_lead;
    dogm_b200::SegPiece* seg_trail;
    int* seg_flags;

    // resampling
    double* cdf;      // N + B
    double* tile_sum; // per CDF tile
    double* tile_off; // prefix of the tile groups before (k_cdf_chain)
    int* res_start;        // [ceil(N / 256)] CDF position of the first offset of every resampling CTA (k_cdf_chain -> k_resample)
    uint32_t* chain_flags; // the tile ticket of k_cdf_chain
    uint32_t chain_epoch, chain_ticket_base;
    int chain_capacity; // CTAs of k_cdf_chain the device holds at once
    unsigned skip_mask;  // DOGM_B200_SKIP (developer ablation)
    unsigned long long* trace_buf; // kTraceSlots start stamps (dogm_trace_arm)
    int n_cdf_tiles;
    int* ancestors; // N

    // noise (injected mode)
    float4* predict_noise;
    float2* birth_noise;
    float2* init_velocity;
    float* resample_u;

    dogm_b200::DeviceScalars* scal;

    // host-side state (dogm.h:193-198)
    bool first_pose_received;
    bool first_measurement_received;
    float position_x, position_y, yaw;
    dogm_b200::CycleShift shift; // pending ego-motion shift: consumed by prediction (particles) and the cell kernel (grid)
    bool shift_particles_pending;
    bool shift_grid_pending;
    uint32_t cycle;

    // dynamic-cell extraction (allocated on first use, grown on demand)
    dogm_dynamic_cell* dyn_buf;
    int dyn_capacity;
    int* dyn_count;
    int* dyn_count_host; // pinned
    // filter registered with dogm_set_dynamic_cell_filter: the cell kernel then appends matching cells straight into a
    // host-mapped pinned buffer, so that the read-out after a cycle costs one small copy and no extra kernel
    bool dyn_filter_on, dyn_list_valid;
    float dyn_filter_occ, dyn_filter_vel;
    int dyn_filter_capacity;
    uint32_t dyn_list_cycle;
    dogm_dynamic_cell* dyn_mapped_host;
    dogm_dynamic_cell* dyn_mapped_dev;
    // early publication of that list: a kernel behind the cell kernel (k_birth_particles) writes {count, sequence number}
    // into host-mapped memory, so dogm_extract_dynamic_cells returns as soon as the list exists - the rest of the cycle
    // (birth particles, CDF, resampling) keeps running while the host consumes the list and enqueues the next scan
    int* dyn_pub_host; // pinned + mapped, 2 ints
    int* dyn_pub_dev;
    int dyn_pub_seq;      // sequence number of the last list whose publication was enqueued
    bool dyn_pub_armed;   // the cell kernel of the running cycle fills the list; the birth kernel behind it publishes it
    bool dyn_pub_pending; // a publication has been enqueued for the current list

    // instrumentation
    bool timer_ready;
    cudaEvent_t timer_e0, timer_e1;
    uint64_t launch_count;
    bool timing;
    std::vector<dogm_b200::TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    double acc_ms[dogm_b200::K_COUNT];
    uint64_t acc_launches[dogm_b200::K_COUNT];
    double last_bytes[dogm_b200::K_COUNT];
    double acc_bytes[dogm_b200::K_COUNT];
};

namespace dogm_b200
{

// instrumentation hooks around every launch
void launch_begin(dogm_handle* h, int id, double algorithmic_bytes);
void launch_end(dogm_handle* h, int id);

// Programmatic dependent launch: every kernel of the library starts with pdl_prologue() and is launched with
// launch_chained(), so the next kernel's CTAs are scheduled into the SM slots the running kernel's tail leaves free and
// wait (griddepcontrol.wait) until the whole predecessor has finished and its writes are visible.  Since every kernel
// waits, completion stays transitive along the stream.
// Optional timeline (dogm_trace_arm): the first CTA of every launch to get past the wait stamps %globaltimer into its
// slot (KernelId * 2 + pass).  With every kernel chained, the gap between two stamps is what the earlier kernel costs
// inside the free-running cycle - tail, flush and launch gap included - without events that would break the chain.
constexpr int kTraceSlots = 2 * K_COUNT;
static __constant__ unsigned long long* c_trace; // one copy per translation unit, bound by trace_bind_*()

// L2 eviction priorities per access (no set-aside, no access-policy window): the 64 MB of particle records are written by the
// prediction and gathered twice later in the cycle (segmented sums, resampling) - they are stored and re-read as evict_last;
// what is written once and not read again inside the cycle (the GridCell array) goes evict_first.
__device__ __forceinline__ uint64_t l2_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint(float4* ptr, const float4 v, const uint64_t policy)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ float4 ld_hint(const float4* ptr, const uint64_t policy)
{
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(ptr), "l"(policy));
    return v;
}

__device__ __forceinline__ void pdl_prologue(int trace_slot)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && c_trace)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(c_trace + trace_slot, t);
    }
}

// developer ablation switch (DOGM_B200_SKIP=<bit mask of KernelId>, effective from cycle 12 on): LaunchScope arms it for
// the launch it brackets; results are garbage, only the timing of the remaining kernels is of interest
inline thread_local bool g_skip_next_launch = false; // (per host thread: band groups launch from several threads)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(cudaStream_t stream, void (*kernel)(KArgs...), long long grid, int block, size_t smem,
                                  Args&&... args)
{
    if (g_skip_next_launch)
    {
        g_skip_next_launch = false;
        return cudaSuccess;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct LaunchScope
{
    dogm_handle* h;
    int id;
    LaunchScope(dogm_handle* h_, int id_, double bytes) : h(h_), id(id_)
    {
        launch_begin(h, id, bytes);
        g_skip_next_launch = h->skip_mask != 0 && h->cycle >= 12 && ((h->skip_mask >> id) & 1u);
    }
    ~LaunchScope() { launch_end(h, id); }
};

// stage launchers (kernels_particles.cu / kernels_cells.cu / kernels_meas.cu); all stream-ordered on h->stream
int ensure_soa(dogm_handle* h); // materialise the SoA block from the records when a caller wants to see it
int run_init_particles(dogm_handle* h);
int run_predict(dogm_handle* h, float dt);
int run_assignment(dogm_handle* h);
int run_occupancy_update(dogm_handle* h, float dt);
int run_persistent_weights(dogm_handle* h, bool defer);
int run_birth(dogm_handle* h);
int run_init_masses(dogm_handle* h); // first-cycle masses + their block scan
int run_init_fill(dogm_handle* h);   // first-cycle particles
int run_born_scan(dogm_handle* h);   // block offsets and total of the born masses
int run_birth_fill(dogm_handle* h, bool fused_scan = false); // birth particles (fused_scan: the kernel scans the block sums itself)
int run_resampling(dogm_handle* h);
int run_cdf(dogm_handle* h);            // joint weight CDF only (first half of run_resampling)
int run_resample_gather(dogm_handle* h); // ancestor search + gather (second half)
bool sort_fused(const dogm_handle* h);
int run_band_outbox(dogm_handle* h); // compacts the particles k_predict flagged into the send boxes, in slot order
int run_band_append(dogm_handle* h, int n_from_lo, int n_from_hi);
// device-paced band cycle (kernels_particles.cu): the three message exchanges and the pull of the neighbours' records
int run_band_publish_sent(dogm_handle* h);
int run_band_collect_sent(dogm_handle* h, int n_pred);
int run_band_pull(dogm_handle* h, int n_pred, const void* outbox_lo, const void* outbox_hi, const float* edge_lo, const float* edge_hi);
int run_band_publish_share(dogm_handle* h, int which); // 0: born mass, 1: joint weight
int run_band_collect_born(dogm_handle* h);
int run_band_collect_weight(dogm_handle* h);
void set_particle_counts(dogm_handle* h, int n, int b); // band mode: current counts and everything derived from them
int materialize_meas(dogm_handle* h);      // kernels_meas.cu: runs a pending cartesian resampling into h->meas
int run_extract_free_mass(dogm_handle* h); // GridCell AoS -> free_cur after dogm_set_grid_cells
int run_init_grid(dogm_handle* h);         // initGridCellsKernel
int run_blocksum_scan(dogm_handle* h, const double* in, double* out_excl, int n, double* total_out, bool publish_dyn = false);
int configure_kernels();
int chain_blocks_per_sm();
int debug_phase_read(int which, void* out_host, size_t bytes); // -DDOGM_PHASE_TRACE builds only
int debug_phase_read_cells(void* out_host, size_t bytes);
int trace_bind_particles(unsigned long long* p);
int trace_bind_cells(unsigned long long* p);
int trace_bind_meas(unsigned long long* p);                   // opt-in shared memory sizes
int run_search_ancestors_f32(dogm_handle* h, const float* d_cdf, int n_cdf, const float* d_draws, int n_draws, int* d_out);
int run_export_noise(dogm_handle* h, uint32_t cycle, float4* d_predict, float2* d_birth, float2* d_init, float* d_resample);
int run_extract_dynamic_cells(dogm_handle* h, float min_occ, float min_vel, dogm_dynamic_cell* d_out, int capacity,
                              int* d_count);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------
// device helpers shared by the kernels
// ---------------------------------------------------------------------------------------------------------
// Cartesian measurement cell from the polar table of a scan: bilinear sample (GL_LINEAR, mip level 0, GL_CLAMP_TO_BORDER with
// border 0, texture.cpp:16-29) at the position cached per cell by k_meas_geom; MeasurementCell {free_mass, occ_mass,
// likelihood, p_A} as in measurement_grid.cu:126-130.  Used by k_meas_apply and by the cell kernel's fused variant.
constexpr int kNoTexel = -(1 << 30);
__device__ __forceinline__ float2 polar_fetch(const float2* __restrict__ table, int K, int H, int bi, int ri)
{
    if (bi < 0 || bi >= K || ri < 0 || ri >= H)
        return make_float2(0.0f, 0.0f);
    return __ldg(table + (size_t)ri * K + bi);
}
__device__ __forceinline__ float4 meas_cell_from_polar(const float4 g, const float2* __restrict__ table, int K, int H)
{
    const int i0 = __float_as_int(g.x), j0 = __float_as_int(g.y);
    float occ_out = 0.0f, free_out = 0.0f;
    if (i0 != kNoTexel)
    {
        const float wu = g.z, wv = g.w;
        const float2 t00 = polar_fetch(table, K, H, i0, j0), t10 = polar_fetch(table, K, H, i0 + 1, j0);
        const float2 t01 = polar_fetch(table, K, H, i0, j0 + 1), t11 = polar_fetch(table, K, H, i0 + 1, j0 + 1);
        const float ob = t00.x + wu * (t10.x - t00.x), ot = t01.x + wu * (t11.x - t01.x);
        const float fb = t00.y + wu * (t10.y - t00.y), ft = t01.y + wu * (t11.y - t01.y);
        occ_out = ob + wv * (ot - ob);
        free_out = fb + wv * (ft - fb);
    }
    return make_float4(free_out, occ_out, 1.0f, 1.0f);
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ unsigned lanemask_le()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// A published sum travels as one 64-bit word: the sums are non-negative, so the sign bit carries the parity of the
// launch epoch.  Every launch publishes every word exactly once, hence a word whose sign bit differs from the epoch's
// parity still holds the previous launch's value.  One relaxed 64-bit load is both the flag and the payload.
__device__ __forceinline__ void publish_f64(double* p, double v, uint32_t epoch)
{
    const unsigned long long bits =
        ((unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull) | ((unsigned long long)(epoch & 1u) << 63);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(bits) : "memory");
}
__device__ __forceinline__ double await_f64(const double* p, uint32_t epoch)
{
    unsigned long long bits;
    for (;;)
    {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(bits) : "l"(p) : "memory");
        if ((uint32_t)(bits >> 63) == (epoch & 1u))
            break;
        __nanosleep(40);
    }
    return __longlong_as_double((long long)(bits & 0x7fffffffffffffffull));
}

// inclusive scan of a double over the 256 threads of a CTA, fixed combination order.
// smem: at least kWarpsPerBlock doubles.  Returns the inclusive prefix; *total receives the CTA sum.
__device__ __forceinline__ double block_inclusive_scan_f64(double v, double* smem, double* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        double t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d)
            v += t;
    }
    if (lane == 31)
        smem[warp] = v;
    __syncthreads();
    double off = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; w++)
    {
        const double s = smem[w];
        if (w < warp)
            off += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return off + v;
}
#endif

} // namespace dogm_b200
