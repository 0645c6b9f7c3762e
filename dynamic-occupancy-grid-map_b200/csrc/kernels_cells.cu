// Cell-side kernels of the DOGM cycle for sm_100a:
//   fused per-cell update (ego-motion grid shift, Dempster-Shafer occupancy update, born/persistent split,
//   persistent-weight normalisers, velocity moments, born-mass block sums), birth-slot distribution and
//   birth-particle initialisation, first-cycle particle initialisation, dynamic-cell extraction.
//
// Reference behaviour restated here (paths relative to the reference's dogm/):
//   src/kernel/ego_motion_compensation.cu:25-43, src/kernel/mass_update.cu:16-93,
//   src/kernel/update_persistent_particles.cu:16-31,59-76, src/kernel/statistical_moments.cu:16-106,
//   src/kernel/init_new_particles.cu:30-195, src/kernel/init.cu:46-93, demo/utils/image_creation.cpp:19-66.
//
// Compiled with -fmad=false: every operator below is one IEEE rounding, in the order written.
#include "dogm_internal.cuh"

namespace dogm_b200
{

// =========================================================================================================
// the fused cell kernel (one thread per cell, 256 cells per CTA)
// =========================================================================================================
struct CellArgs
{
    int C, gs;
    int* cell_start; // read, then reset to -1 for the next cycle's assignment (reinitGridParticleIndices, init.cu:86-93)
    const int* cell_end;
    const CellSums* sums;
    const dogm_meas_cell* meas;  // the measurement grid of this cycle (may be the caller's device buffer)
    dogm_meas_cell* meas_copy;   // when non-null: the handle's own copy is written on the way (dogm.cu:207-208)
    const float* free_cur;
    float* free_next;
    dogm_grid_cell* grid;
    float* born_masses;
    float4* coef;
    double* blk_sum;
    double* prefix; // block-local inclusive prefix of the born masses (double), for the birth slot distribution
    // optional fused extraction of dynamic cells (computeCellsWithVelocity, demo/utils/image_creation.cpp:19-66)
    dogm_dynamic_cell* dyn_out; // device-visible (host-mapped) record buffer, or null
    int* dyn_count;
    int dyn_capacity;
    float dyn_min_occ, dyn_min_vel;
    float p_B, alpha;
    int shift_active, x_move, y_move;
    // band mode: this handle holds rows [row0, row0 + rows) of a grid of gs rows; rows that an ego-motion shift pulls in
    // from a neighbour band are read from the halo copies (halo_lo: the halo_rows rows below row0, halo_hi: above)
    int row0, rows, halo_rows;
    const float* halo_lo;
    const float* halo_hi;
    // kLazyMeas: the measurement cell is sampled from the polar table of the scan here (dogm_meas_generate_into) instead of
    // being read from `meas`; `meas_copy` receives it
    const float4* geom;
    const float2* polar;
    int polar_K, polar_H;
    // shortcut for blocks in which nothing happens (dogm_handle::blk_age): ages of the previous cycle (read) and of this one (written)
    const uint8_t* age_prev;
    uint8_t* age_next;
};

#ifndef DOGM_CELL_MINBLOCKS
#define DOGM_CELL_MINBLOCKS 8
#endif
#ifndef DOGM_CELL_MINBLOCKS_LAZY
#define DOGM_CELL_MINBLOCKS_LAZY 8
#endif
// kQuiet: with the shortcut for blocks in which nothing happens (large grids; on a small grid it only costs its bookkeeping)
template <bool kLazyMeas, bool kQuiet>
__global__ void __launch_bounds__(kCellBlock, kLazyMeas ? DOGM_CELL_MINBLOCKS_LAZY : DOGM_CELL_MINBLOCKS) k_cell(CellArgs a)
{
    pdl_prologue(K_CELL * 2);
    __shared__ double s_scan[kWarpsPerBlock];
    // staging for the 64-byte GridCell records: [warp][part][cell], row stride 34 float4 keeps both the per-thread
    // writes and the transposed reads free of bank conflicts
    __shared__ float4 s_out[kWarpsPerBlock][4][34];
    const int c = blockIdx.x * kCellBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = c < a.C;
    float rho_b = 0.0f;
    bool dyn_hit = false;
    dogm_dynamic_cell dyn_rec;
    // all loads that do not depend on the cell being occupied go out together, ahead of the first store
    float4 zq = make_float4(0.0f, 0.0f, 1.0f, 1.0f);
    int start = -1;
    float free_prev = 0.0f;
    const int age = kQuiet ? a.age_prev[blockIdx.x] : 0;
    if (valid)
    {
        start = a.cell_start[c];
        if constexpr (kLazyMeas)
            zq = meas_cell_from_polar(__ldcs(a.geom + c), a.polar, a.polar_K, a.polar_H);
        else
            zq = __ldg(reinterpret_cast<const float4*>(a.meas + c));
        // ego-motion compensation of the grid (updatePose dogm.cu:175-193, moveMapKernel
        // ego_motion_compensation.cu:25-43): of all cell fields only free_mass survives into the next cycle,
        // so the shift is a shifted read of the previous free masses; vacated cells read 0 (dogm.cu:185)
        if (a.shift_active)
        {
            const int x = c % a.gs, y = c / a.gs + a.row0;
            const int nx = x + a.x_move, ny = y + a.y_move;
            if (nx > 0 && nx < a.gs && ny > 0 && ny < a.gs)
            {
                const int nl = ny - a.row0; // row inside this band
                if (nl >= 0 && nl < a.rows)
                    free_prev = __ldg(a.free_cur + nx + a.gs * nl);
                else if (nl < 0 && a.halo_lo && -nl <= a.halo_rows)
                    free_prev = __ldg(a.halo_lo + nx + a.gs * (a.halo_rows + nl));
                else if (nl >= a.rows && a.halo_hi && nl - a.rows < a.halo_rows)
                    free_prev = __ldg(a.halo_hi + nx + a.gs * (nl - a.rows));
            }
        }
        else
        {
            free_prev = __ldg(a.free_cur + c);
        }
    }
    const bool z_none = zq.x == 0.0f && zq.y == 0.0f && zq.z == 1.0f && zq.w == 1.0f;
    // a cell in which nothing happens: no particles, no measurement, no free mass left - its outputs are the empty record
    const bool cell_idle = start < 0 && free_prev == 0.0f && z_none;
    // Shortcut: the whole block has been like that for the last two cycles (both free-mass buffers, the grid cells, the born
    // masses, the prefix and the handle's measurement copy already hold what would be written) and is again: nothing to
    // compute, nothing to store.  (Most of a large grid lies outside the sensor's field of view.)
    if (kQuiet && age >= 2) // (block-uniform)
    {
        if (__syncthreads_and(cell_idle))
        {
            if (threadIdx.x == 0)
                a.age_next[blockIdx.x] = 2;
            return;
        }
    }
    if (valid)
    {
        const bool occupied = start >= 0;
        int end = -1;
        CellSums cs;
        cs.s0 = cs.s1 = cs.s2 = cs.s3 = cs.s4 = cs.s5 = 0.0f;
        if (occupied)
        {
            end = __ldg(a.cell_end + c);
            const float4* sp = reinterpret_cast<const float4*>(a.sums + c);
            const float4 lo = __ldg(sp), hi = __ldg(sp + 1);
            cs.s0 = lo.x;
            cs.s1 = lo.y;
            cs.s2 = lo.z;
            cs.s3 = lo.w;
            cs.s4 = hi.x;
            cs.s5 = hi.y;
            a.cell_start[c] = -1;
        }
        if (a.meas_copy)
            *reinterpret_cast<float4*>(a.meas_copy + c) = zq;
        const float z_free = zq.x, z_occ = zq.y, lik = zq.z, p_A = zq.w;

        // gridCellPredictionUpdateKernel, mass_update.cu:61-93
        float m_occ_pred = occupied ? cs.s0 : 0.0f;
        float over_div = 0.0f;
        if (m_occ_pred > 1.0f)
        { // normalize_weights, mass_update.cu:51-59: applied per particle by k_weights
            over_div = m_occ_pred;
            m_occ_pred = 1.0f;
        }
        const float fa = a.alpha * free_prev;
        const float fb = 1.0f - m_occ_pred;
        const float m_free_pred = fminf(fa, fb);
        const float unknown_pred = 1.0f - m_occ_pred - m_free_pred;
        const float meas_unknown = 1.0f - z_free - z_occ;
        const float K = m_free_pred * z_occ + m_occ_pred * z_free;
        const float occ_up = (m_occ_pred * meas_unknown + unknown_pred * z_occ + m_occ_pred * z_occ) / (1.0f - K);
        const float free_up = (m_free_pred * meas_unknown + unknown_pred * z_free + m_free_pred * z_free) / (1.0f - K);
        rho_b = (occ_up * a.p_B * (1.0f - m_occ_pred)) / (m_occ_pred + a.p_B * (1.0f - m_occ_pred));
        const float rho_p = occ_up - rho_b;

        // updatePersistentParticlesKernel2, update_persistent_particles.cu:59-76 (sum of likelihood-weighted
        // weights of the cell = likelihood * predicted occupancy: the likelihood is a per-cell constant)
        float mu_A = 0.0f, mu_UA = 0.0f;
        float mean_x = 0.0f, mean_y = 0.0f, var_x = 0.0f, var_y = 0.0f, covar = 0.0f;
        if (occupied)
        {
            const float m_occ_accum = lik * m_occ_pred;
            mu_A = m_occ_accum > 0.0f ? rho_p / m_occ_accum : 0.0f;
            mu_UA = m_occ_pred > 0.0f ? rho_p / m_occ_pred : 0.0f;
            const float cA = p_A * mu_A;
            const float cUA = (1.0f - p_A) * mu_UA;
            a.coef[c] = make_float4(lik, cA, cUA, over_div);

            // statisticalMomentsKernel2, statistical_moments.cu:78-106: the updated weight of every particle of
            // the cell is its predicted weight times one per-cell factor, so the moment sums of the predicted
            // weights (k_segsum) are rescaled instead of summed again
            if (rho_p > 0.0f)
            {
                float kfac = cA * lik + cUA;
                if (over_div > 0.0f)
                    kfac = kfac / over_div;
                const float inv = 1.0f / rho_p;
                mean_x = inv * (kfac * cs.s1);
                mean_y = inv * (kfac * cs.s2);
                var_x = inv * (kfac * cs.s3) - mean_x * mean_x;
                var_y = inv * (kfac * cs.s4) - mean_y * mean_y;
                covar = inv * (kfac * cs.s5) - mean_x * mean_y;
            }
        }

        a.born_masses[c] = rho_b;
        a.free_next[c] = free_up;
        if (a.dyn_out)
        {
            const float occ = occ_up + 0.5f * (1.0f - occ_up - free_up);
            const float det = var_x * var_y - covar * covar;
            const float maha = (var_y * mean_x * mean_x - 2.0f * covar * mean_x * mean_y + var_x * mean_y * mean_y) / det;
            if (occ >= a.dyn_min_occ && maha >= a.dyn_min_vel)
            {
                dyn_hit = true;
                dyn_rec.cell_idx = c;
                dyn_rec.occupancy = occ;
                dyn_rec.mean_x_vel = mean_x;
                dyn_rec.mean_y_vel = mean_y;
                dyn_rec.var_x_vel = var_x;
                dyn_rec.var_y_vel = var_y;
                dyn_rec.covar_xy_vel = covar;
                dyn_rec.mahalanobis = maha;
            }
        }
        s_out[warp][0][lane] = make_float4(__int_as_float(start), __int_as_float(end), rho_b, rho_p);
        s_out[warp][1][lane] = make_float4(free_up, occ_up, m_occ_pred, mu_A);
        s_out[warp][2][lane] = make_float4(mu_UA, 0.0f, 0.0f, mean_x); // w_A / w_UA: k_birth_cells, for cells that own slots
        s_out[warp][3][lane] = make_float4(mean_y, var_x, var_y, covar);
    }
    if (a.dyn_out)
    { // warp-aggregated append of the cells that pass the filter
        const unsigned m = __ballot_sync(0xffffffffu, dyn_hit);
        if (m)
        {
            int base = 0;
            if (lane == __ffs(m) - 1)
                base = atomicAdd(a.dyn_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (dyn_hit)
            {
                const int at = base + __popc(m & lanemask_lt());
                if (at < a.dyn_capacity)
                { // two 16-byte stores (the buffer may be host-mapped: every store instruction is a PCIe write)
                    float4* o = reinterpret_cast<float4*>(a.dyn_out + at);
                    o[0] = make_float4(__int_as_float(dyn_rec.cell_idx), dyn_rec.occupancy, dyn_rec.mean_x_vel, dyn_rec.mean_y_vel);
                    o[1] = make_float4(dyn_rec.var_x_vel, dyn_rec.var_y_vel, dyn_rec.covar_xy_vel, dyn_rec.mahalanobis);
                }
            }
        }
    }
    // the warp's 32 GridCells are 2 KB of contiguous memory: write them as four fully coalesced 512-byte rows
    __syncwarp();
    {
        const int warp_cell0 = blockIdx.x * kCellBlock + warp * 32;
        float4* gw = reinterpret_cast<float4*>(a.grid + warp_cell0);
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const int f = j * 32 + lane; // float4 index inside the warp's block: cell f/4, part f%4
            if (warp_cell0 + (f >> 2) < a.C)
            {
#ifdef DOGM_NO_L2_HINTS
                gw[f] = s_out[warp][f & 3][f >> 2];
#else
                st_hint(gw + f, s_out[warp][f & 3][f >> 2], l2_evict_first()); // (nobody reads the cells again inside the cycle)
#endif
            }
        }
    }
    double total;
    const double incl = block_inclusive_scan_f64(valid ? (double)rho_b : 0.0, s_scan, &total);
    if (valid)
        a.prefix[c] = incl;
    if (kQuiet)
    {
        const int idle = __syncthreads_and(cell_idle);
        if (threadIdx.x == 0)
            a.age_next[blockIdx.x] = idle ? (uint8_t)(age >= 1 ? 2 : 1) : (uint8_t)0;
    }
    if (threadIdx.x == 0)
        a.blk_sum[blockIdx.x] = total;
}

// =========================================================================================================
// slot distribution: accumulate + normalize_particle_orders + calc_start_idx/calc_end_idx
// (init_new_particles.cu:30-43,66-74).  The running sum over all cells is  grp_off[block / 256] + blk_off[block] +
// prefix[cell]  (doubles: block-local prefix from the cell kernel; the block sums are scanned in groups of 256 blocks,
// blk_off being the prefix inside the group and grp_off the sum of the groups in front); the exclusive end slot of a cell,
// int(float(sum) * (count / total)), is evaluated where it is needed instead of being stored per cell.
// Two producers of (grp_off, blk_off): the birth kernel itself (run_birth_fill: its first CTAs scan one group each, every CTA
// then adds up the few group sums in shared memory - no launch between the cell kernel and the birth kernel), or
// k_blocksum_scan (band mode and the first cycle, where a global normaliser has to leave the device in between): it writes
// the complete offsets into blk_off, and grp_off points at zeros.
// =========================================================================================================
constexpr int kBlkGroup = 256; // 256-cell blocks per scan group

struct SlotView
{
    const double* grp_off;
    const double* blk_off;
    const double* blk_sum;
    const double* prefix;
    int n_blocks, C;
    float scale; // (float)count / (float)total
    // band mode: mass of the bands in front of this one and the global number of this band's first slot (0, 0 otherwise)
    double base_off;
    int slot_base;
};

__device__ __forceinline__ double block_offset(const SlotView& v, int b)
{ // (blk_off may have been written by other CTAs of the running kernel: read through L2)
    return v.grp_off[b / kBlkGroup] + __ldcg(v.blk_off + b);
}
__device__ __forceinline__ int slot_end_of_block(const SlotView& v, int b)
{
    return __float2int_rz((float)((v.base_off + block_offset(v, b)) + v.blk_sum[b]) * v.scale) - v.slot_base;
}
__device__ __forceinline__ int slot_end_of_cell(const SlotView& v, int c)
{
    return __float2int_rz((float)((v.base_off + block_offset(v, c / kCellBlock)) + v.prefix[c]) * v.scale) - v.slot_base;
}
__device__ __forceinline__ int slot_start_of_cell(const SlotView& v, int c)
{
    if (c == 0)
        return 0;
    // the cell in front of the first cell of a block ends where that block's offset puts it
    return (c % kCellBlock) ? slot_end_of_cell(v, c - 1)
                            : __float2int_rz((float)(v.base_off + block_offset(v, c / kCellBlock)) * v.scale) - v.slot_base;
}

// owner cell of slot s: first cell j whose end slot exceeds s (two-level search: 256-cell blocks, then cells)
__device__ __forceinline__ int find_slot_owner(const SlotView& v, int s)
{
    int lo = 0, hi = v.n_blocks;
    while (lo < hi)
    {
        const int mid = lo + ((hi - lo) >> 1);
        if (slot_end_of_block(v, mid) > s)
            hi = mid;
        else
            lo = mid + 1;
    }
    if (lo >= v.n_blocks)
        return -1;
    int clo = lo * kCellBlock, chi = min(v.C, clo + kCellBlock);
    while (clo < chi)
    {
        const int mid = clo + ((chi - clo) >> 1);
        if (slot_end_of_cell(v, mid) > s)
            chi = mid;
        else
            clo = mid + 1;
    }
    return clo < v.C ? clo : -1;
}

// per-cell part of initNewParticlesKernel1 (init_new_particles.cu:134-143)
struct BirthWeights
{
    int start, nu_A;
    float w_A, w_UA;
};
__device__ __forceinline__ BirthWeights birth_weights_of_cell(const SlotView& v, int j, float p_A, float born_mass)
{
    BirthWeights r;
    r.start = slot_start_of_cell(v, j);
    const int e = slot_end_of_cell(v, j);
    const int num_new = e > r.start ? e - r.start : 0;
    r.nu_A = __float2int_rz(roundf((float)num_new * p_A));
    const int nu_UA = num_new - r.nu_A;
    r.w_A = r.nu_A > 0 ? (p_A * born_mass) / (float)r.nu_A : 0.0f;
    r.w_UA = nu_UA > 0 ? ((1.0f - p_A) * born_mass) / (float)nu_UA : 0.0f;
    return r;
}

struct BirthArgs
{
    int B, gs;
    int B_glob;    // birth particles of the whole grid (the slot scale); == B without bands
    int cell_base; // number of this band's first cell in the whole grid (0 without bands)
    SlotView slots;
    const DeviceScalars* scal;
    const float* born_masses;
    const dogm_meas_cell* meas;
    dogm_grid_cell* grid;
    ParticleSet birth;
    const float2* noise;
    int noise_injected;
    float stddev_velocity;
    uint64_t seed;
    uint32_t cycle;
    // early publication of the dynamic-cell list (see dogm_handle::dyn_pub_host): written by one thread of this kernel, where
    // the system-wide fence hides behind the other CTAs' work (in the single-CTA scan in front it would delay the whole cycle)
    const int* pub_count;
    int* pub_host;
    int pub_seq;
    // fused scan of the block sums (see SlotView): group g is scanned by CTA g, group sums travel as parity words
    int n_groups;          // 0: the offsets were produced by k_blocksum_scan
    double* blk_off_w;     // == slots.blk_off
    double* grp_word;      // [n_groups]
    uint32_t epoch;
    double* total_out;     // DeviceScalars::born_total
    const BandCounts* cnt; // device-paced band cycle: this band's birth count and slot base live on the device
};

// Developer build (-DDOGM_PHASE_TRACE): phase stamps of k_birth_particles (thread 0 of every CTA), see tools/phase_trace.py
#ifdef DOGM_PHASE_TRACE
constexpr int kPhaseCtasC = 4096, kPhaseSlotsC = 8;
__device__ unsigned long long g_phase_cells[kPhaseCtasC][kPhaseSlotsC];
__device__ __forceinline__ void phase_stamp_c(unsigned cta, int slot, unsigned dep)
{
    if (threadIdx.x == 0 && cta < (unsigned)kPhaseCtasC)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer; // %1" : "=l"(t) : "r"(dep) : "memory");
        g_phase_cells[cta][slot] = t;
    }
}
#define PHASE_STAMP_C(cta, slot, dep) phase_stamp_c(cta, slot, (unsigned)(dep))
#else
#define PHASE_STAMP_C(cta, slot, dep)
#endif
int debug_phase_read_cells(void* out_host, size_t bytes)
{
#ifdef DOGM_PHASE_TRACE
    if (bytes > sizeof(g_phase_cells))
        return DOGM_ERR_INVALID_ARGUMENT;
    return (int)cudaMemcpyFromSymbol(out_host, g_phase_cells, bytes, 0, cudaMemcpyDeviceToHost);
#else
    (void)out_host, (void)bytes;
    return DOGM_ERR_UNSUPPORTED;
#endif
}

// initBirthParticlesKernel (init.cu:46-67) + initNewParticlesKernel1 (init_new_particles.cu:126-155, with the
// deterministic ownership rule "slot s belongs to cell j iff start_j <= s <= end_j") + initNewParticlesKernel2 (:157-195)
__device__ __forceinline__ void birth_block(const BirthArgs& a, const int vblock, const int n_birth, double* s_grp)
{
    __shared__ double s_scan[kWarpsPerBlock];
    const int s = vblock * kBlock + threadIdx.x;
    if (a.pub_host && s == 0)
    {
        const int found = __ldcg(a.pub_count);
        *reinterpret_cast<volatile int*>(a.pub_host) = found;
        __threadfence_system();
        *reinterpret_cast<volatile int*>(a.pub_host + 1) = a.pub_seq;
    }
    // the velocity draw does not depend on the slot distribution: it overlaps the ticket and the group scans
    float2 vel = make_float2(0.0f, 0.0f);
    if (s < n_birth)
    {
        if (a.noise_injected)
            vel = a.noise[s];
        else
        {
            const float4 g = philox_normal4(a.seed, (uint32_t)s, STAGE_BIRTH, a.cycle);
            vel = make_float2(g.x * a.stddev_velocity, g.y * a.stddev_velocity);
        }
    }
    PHASE_STAMP_C(blockIdx.x, 0, 0);
    SlotView v = a.slots;
    if (a.cnt)
    {
        v.base_off = a.cnt->born_base;
        v.slot_base = a.cnt->birth_slot_base;
    }
    double born_total;
    if (a.n_groups > 0)
    {
        // The first CTAs of the grid scan one group of 256 block sums each (fixed order inside a group) and publish the group
        // sum; they wait for nobody and are dispatched before every CTA behind them, so every wait below ends.  Every CTA then
        // adds up the group sums in group order.
        for (int g = vblock; g < a.n_groups; g += (int)gridDim.x)
        {
            const int j = g * kBlkGroup + (int)threadIdx.x;
            const double val = j < v.n_blocks ? v.blk_sum[j] : 0.0;
            double total;
            const double incl = block_inclusive_scan_f64(val, s_scan, &total);
            if (j < v.n_blocks)
                __stcg(a.blk_off_w + j, incl - val);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0)
                publish_f64(a.grp_word + g, total, a.epoch);
        }
        for (int g = threadIdx.x; g < a.n_groups; g += kBlock)
            s_grp[g] = await_f64(a.grp_word + g, a.epoch);
        __threadfence(); // the group's block offsets were written before its word
        __syncthreads();
        if (a.n_groups <= kBlock)
        {
            const double val = (int)threadIdx.x < a.n_groups ? s_grp[threadIdx.x] : 0.0;
            double total;
            const double incl = block_inclusive_scan_f64(val, s_scan, &total);
            if ((int)threadIdx.x < a.n_groups)
                s_grp[threadIdx.x] = incl - val;
            if (threadIdx.x == 0)
                s_grp[a.n_groups] = total;
        }
        else
        { // very large grids: consecutive groups per thread
            const int per = (a.n_groups + kBlock - 1) / kBlock;
            const int g0 = min((int)threadIdx.x * per, a.n_groups), g1 = min(g0 + per, a.n_groups);
            double local = 0.0;
            for (int g = g0; g < g1; g++)
                local += s_grp[g];
            double total;
            double run = block_inclusive_scan_f64(local, s_scan, &total) - local;
            for (int g = g0; g < g1; g++)
            {
                const double val = s_grp[g];
                s_grp[g] = run;
                run += val;
            }
            if (threadIdx.x == 0)
                s_grp[a.n_groups] = total;
        }
        __syncthreads();
        v.grp_off = s_grp;
        born_total = s_grp[a.n_groups];
        if (vblock == 0 && threadIdx.x == 0)
            *a.total_out = born_total;
    }
    else
        born_total = a.scal->born_total;
    if (s >= n_birth)
        return;
    v.scale = (float)a.B_glob / (float)born_total;
    PHASE_STAMP_C(blockIdx.x, 1, __float_as_uint(v.scale));
    int j = find_slot_owner(v, s);
    PHASE_STAMP_C(blockIdx.x, 2, j);
    bool assoc = false;
    float weight;
    if (j >= 0)
    {
        const BirthWeights bw = birth_weights_of_cell(v, j, a.meas[j].p_A, a.born_masses[j]);
        assoc = s <= bw.start + bw.nu_A;
        weight = assoc ? bw.w_A : bw.w_UA;
        if (s == bw.start)
        { // the cell's first slot also records the weights in the grid cell (store_weights, :60-64)
            a.grid[j].w_A = bw.w_A;
            a.grid[j].w_UA = bw.w_UA;
        }
    }
    else
    { // a slot no cell owns keeps its previous cell index, as in the reference
        j = a.birth.idx[s];
        weight = birth_weights_of_cell(v, j, a.meas[j].p_A, a.born_masses[j]).w_UA;
    }
    PHASE_STAMP_C(blockIdx.x, 3, __float_as_uint(weight) ^ __float_as_uint(vel.x));
    const float x = (float)(j % a.gs) + 0.5f;
    const float y = (float)(j + a.cell_base) / (float)a.gs + 0.5f; // float division, init_new_particles.cu:173
    a.birth.idx[s] = j;
    a.birth.assoc[s] = assoc ? 1 : 0;
    a.birth.weight[s] = weight;
    a.birth.state[s] = make_float4(x, y, vel.x, vel.y);
    PHASE_STAMP_C(blockIdx.x, 4, 0);
}

__global__ void __launch_bounds__(kBlock) k_birth_particles(const BirthArgs a)
{
    pdl_prologue(K_BIRTH_PARTICLES * 2);
    extern __shared__ double s_grp[]; // [n_groups + 1] exclusive prefix of the group sums; the last entry is the total
    if (!a.cnt)
    {
        birth_block(a, (int)blockIdx.x, a.B, s_grp);
        return;
    }
    // device-paced band cycle: the band's birth count is read from the device, the grid is the host's estimate (no fused scan here)
    const int n_birth = a.cnt->B;
    for (int b = blockIdx.x; b * kBlock < n_birth || b == 0; b += gridDim.x)
        birth_block(a, b, n_birth, s_grp);
}

// =========================================================================================================
// first-cycle initialisation: copyMassesKernel + initParticlesKernel1/2 (init_new_particles.cu:76-124)
// =========================================================================================================
__global__ void __launch_bounds__(kCellBlock) k_init_masses(const dogm_meas_cell* __restrict__ meas, float* masses, int C,
                                                            double* blk_sum, double* prefix)
{
    pdl_prologue(K_INIT_MASSES * 2);
    __shared__ double s_scan[kWarpsPerBlock];
    const int c = blockIdx.x * kCellBlock + threadIdx.x;
    float m = 0.0f;
    if (c < C)
    {
        m = meas[c].occ_mass;
        masses[c] = m;
    }
    double total;
    const double incl = block_inclusive_scan_f64((double)m, s_scan, &total);
    if (c < C)
        prefix[c] = incl;
    if (threadIdx.x == 0)
        blk_sum[blockIdx.x] = total;
}

struct InitArgs
{
    int N, gs;
    int N_glob;    // particles of the whole grid (slot scale and initial weight); == N without bands
    int row0;      // first row of this band (0 without bands)
    SlotView slots;
    const DeviceScalars* scal;
    ParticleSet p;
    const float2* init_velocity;
    int noise_injected;
    float init_max_velocity;
    uint64_t seed;
    uint32_t cycle;
};

__global__ void __launch_bounds__(kBlock) k_init_particles(InitArgs a)
{
    pdl_prologue(K_INIT_PARTICLES * 2);
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= a.N)
        return;
    SlotView sv = a.slots;
    sv.scale = (float)a.N_glob / (float)a.scal->born_total;
    int j = find_slot_owner(sv, i);
    if (j < 0)
        j = a.p.idx[i];
    float2 v;
    if (a.noise_injected)
        v = a.init_velocity[i];
    else
    { // curand_uniform(state, -v, v) = min + (max - min) * (1 - u), u in (0,1]  (cuda_utils.h:31-35)
        const Philox4 r = philox4x32_10((uint32_t)i, STAGE_INIT, a.cycle, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const float lo = -a.init_max_velocity;
        const float span = a.init_max_velocity - lo;
        v = make_float2(lo + span * (1.0f - u01_open_low(r.x)), lo + span * (1.0f - u01_open_low(r.y)));
    }
    const float x = (float)(j % a.gs) + 0.5f;
    const float y = (float)(j / a.gs + a.row0) + 0.5f; // integer division, init_new_particles.cu:115
    a.p.idx[i] = j;
    a.p.weight[i] = 1.0f / (float)a.N_glob;
    a.p.state[i] = make_float4(x, y, v.x, v.y);
}

// =========================================================================================================
// small utilities
// =========================================================================================================
__global__ void __launch_bounds__(kBlock) k_extract_free(const dogm_grid_cell* __restrict__ grid, float* free_cur, int C)
{
    pdl_prologue(K_MISC * 2);
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c < C)
        free_cur[c] = grid[c].free_mass;
}

__global__ void __launch_bounds__(kBlock) k_init_grid(dogm_grid_cell* grid, dogm_meas_cell* meas, float* free_a,
                                                      float* free_b, int* cell_start, int C)
{ // initGridCellsKernel, init.cu:69-84 (all other GridCell fields start at 0 here; the reference leaves them uninitialised)
    pdl_prologue(K_MISC * 2);
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= C)
        return;
    float4* g = reinterpret_cast<float4*>(grid + c);
    g[0] = make_float4(__int_as_float(-1), __int_as_float(-1), 0.f, 0.f);
    g[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(meas + c) = make_float4(0.f, 0.f, 1.f, 1.f); // free, occ, likelihood, p_A
    free_a[c] = 0.f;
    free_b[c] = 0.f;
    cell_start[c] = -1;
}

// computeCellsWithVelocity, demo/utils/image_creation.cpp:19-66, as a stream compaction on the device
__global__ void __launch_bounds__(kBlock) k_extract_dynamic(const dogm_grid_cell* __restrict__ grid, int C, float min_occ,
                                                            float min_vel, dogm_dynamic_cell* out, int capacity,
                                                            int* count)
{
    pdl_prologue(K_MISC * 2);
    const int c = blockIdx.x * kBlock + threadIdx.x;
    bool hit = false;
    dogm_dynamic_cell rec;
    if (c < C)
    {
        const float4* g = reinterpret_cast<const float4*>(grid + c);
        const float4 q1 = g[1], q2 = g[2], q3 = g[3];
        const float free_mass = q1.x, occ_mass = q1.y;
        const float mx = q2.w, my = q3.x, vxx = q3.y, vyy = q3.z, cxy = q3.w;
        const float occ = occ_mass + 0.5f * (1.0f - occ_mass - free_mass);
        const float det = vxx * vyy - cxy * cxy;
        const float maha = (vyy * mx * mx - 2.0f * cxy * mx * my + vxx * my * my) / det;
        hit = occ >= min_occ && maha >= min_vel;
        rec.cell_idx = c;
        rec.occupancy = occ;
        rec.mean_x_vel = mx;
        rec.mean_y_vel = my;
        rec.var_x_vel = vxx;
        rec.var_y_vel = vyy;
        rec.covar_xy_vel = cxy;
        rec.mahalanobis = maha;
    }
    // warp-aggregated append
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m)
    {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(m) - 1)
            base = atomicAdd(count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (hit)
        {
            const int at = base + __popc(m & lanemask_lt());
            if (at < capacity)
                out[at] = rec;
        }
    }
}

// =========================================================================================================
// host-side launchers
// =========================================================================================================
static SlotView make_slot_view(dogm_handle* h)
{
    SlotView v;
    v.grp_off = h->grp_zero;
    v.blk_off = h->blk_off;
    v.blk_sum = h->blk_sum;
    v.prefix = h->cell_prefix;
    v.n_blocks = h->n_cell_blocks;
    v.C = h->C;
    v.scale = 0.0f;
    v.base_off = h->band.enabled ? h->band.born_base : 0.0;
    v.slot_base = h->band.enabled ? h->band.birth_slot_base : 0;
    return v;
}

// first-cycle initialisation in two steps, so that a band orchestrator can put the global normaliser in between
int run_init_masses(dogm_handle* h)
{
    {
        LaunchScope ls(h, K_INIT_MASSES, 8.0 * h->C);
        launch_chained(h->stream, k_init_masses, h->n_cell_blocks, kCellBlock, 0, h->meas, h->born_masses, h->C, h->blk_sum,
                                                                      h->cell_prefix);
    }
    return run_blocksum_scan(h, h->blk_sum, h->blk_off, h->n_cell_blocks, &h->scal->born_total);
}

int run_init_fill(dogm_handle* h)
{
    if (h->N <= 0)
        return 0;
    InitArgs a;
    a.N = h->N;
    a.gs = h->gs;
    a.N_glob = h->band.enabled ? h->band.n_glob : h->N;
    a.row0 = h->band.row0;
    a.slots = make_slot_view(h);
    a.scal = h->scal;
    a.p = h->pa;
    a.init_velocity = h->init_velocity;
    a.noise_injected = (h->opts.noise_mode == DOGM_NOISE_INJECTED) ? 1 : 0;
    a.init_max_velocity = h->params.init_max_velocity;
    a.seed = h->opts.seed + h->band.salt;
    a.cycle = h->cycle;
    {
        LaunchScope ls(h, K_INIT_PARTICLES, 24.0 * h->N);
        launch_chained(h->stream, k_init_particles, div_up(h->N, kBlock), kBlock, 0, a);
    }
    h->hist0_valid = false;
    h->pa_current = true;
    h->rec_valid = false;
    h->sorted_valid = false;
    return (int)cudaGetLastError();
}

int run_init_particles(dogm_handle* h)
{
    int e = run_init_masses(h);
    return e ? e : run_init_fill(h);
}

int run_occupancy_update(dogm_handle* h, float dt)
{
    if (h->grid_swap_pending)
    { // pipelined read-out: the cells of the previous cycle are still on their way to the host; this cycle writes the other
      // buffer (every field of every cell is written by the dense cell kernel, nothing of the old buffer is read), once the copy
      // that last read THAT buffer has finished
        h->grid_swap_pending = false;
        dogm_grid_cell* t = h->grid;
        h->grid = h->grid_alt;
        h->grid_alt = t;
        const cudaEvent_t e = h->grid_copy_ev[0];
        h->grid_copy_ev[0] = h->grid_copy_ev[1];
        h->grid_copy_ev[1] = e;
        const bool b = h->grid_copy_busy[0];
        h->grid_copy_busy[0] = h->grid_copy_busy[1];
        h->grid_copy_busy[1] = b;
        if (h->grid_copy_busy[0])
        {
            cudaStreamWaitEvent(h->stream, h->grid_copy_ev[0], 0);
            h->grid_copy_busy[0] = false;
        }
    }
    CellArgs a;
    a.C = h->C;
    a.gs = h->gs;
    a.cell_start = h->cell_start;
    a.cell_end = h->cell_end;
    a.sums = h->cell_sums;
    a.meas = h->meas_src ? h->meas_src : h->meas;
    a.meas_copy = h->meas_src ? h->meas : nullptr;
    a.free_cur = h->free_cur;
    a.free_next = h->free_next;
    a.grid = h->grid;
    a.born_masses = h->born_masses;
    a.coef = h->cell_coef;
    a.blk_sum = h->blk_sum;
    a.prefix = h->cell_prefix;
    a.p_B = h->params.birth_prob;
    a.alpha = powf(h->params.freespace_discount, dt); // std::pow(float, float), dogm.cu:291
    a.shift_active = (h->shift_grid_pending && h->shift.active) ? 1 : 0;
    a.x_move = h->shift.x_move;
    a.y_move = h->shift.y_move;
    a.row0 = h->band.row0;
    a.rows = h->band.rows;
    a.halo_rows = h->band.halo_rows;
    a.halo_lo = h->band.halo_valid ? h->band.halo[0] : nullptr;
    a.halo_hi = h->band.halo_valid ? h->band.halo[1] : nullptr;
    a.dyn_out = nullptr;
    a.dyn_count = h->dyn_count;
    a.dyn_capacity = h->dyn_filter_capacity;
    a.dyn_min_occ = h->dyn_filter_occ;
    a.dyn_min_vel = h->dyn_filter_vel;
    if (h->dyn_filter_on)
    {
        a.dyn_out = h->dyn_mapped_dev;
        if (h->N <= 0) // otherwise k_segfix cleared the counter
            cudaMemsetAsync(h->dyn_count, 0, sizeof(int), h->stream);
        h->dyn_list_cycle = h->cycle;
        h->dyn_list_valid = true;
        h->dyn_pub_armed = true;
        h->dyn_pub_pending = false;
    }
    a.geom = nullptr;
    a.polar = nullptr;
    a.polar_K = a.polar_H = 0;
    a.age_prev = h->blk_age[0];
    a.age_next = h->blk_age[1];
    {
        uint8_t* t = h->blk_age[0];
        h->blk_age[0] = h->blk_age[1];
        h->blk_age[1] = t;
    }
    const bool lazy = h->lazy_meas.pending && !h->meas_src;
    if (lazy)
    {
        a.geom = h->lazy_meas.geom;
        a.polar = h->lazy_meas.polar;
        a.polar_K = h->lazy_meas.K;
        a.polar_H = h->lazy_meas.H;
        a.meas = nullptr;
        a.meas_copy = h->meas;
    }
    h->lazy_meas.pending = false; // consumed, or superseded by the caller's grid
    h->cell_kernel_done = false;
    {
        LaunchScope ls(h, K_CELL, 96.0 * h->C);
        const bool quiet = !h->quiet_off && !h->grid_alt; // (the shortcut relies on the ONE grid buffer holding last cycle's cells)
        if (lazy)
        {
            if (quiet)
                launch_chained(h->stream, k_cell<true, true>, h->n_cell_blocks, kCellBlock, 0, a);
            else
                launch_chained(h->stream, k_cell<true, false>, h->n_cell_blocks, kCellBlock, 0, a);
        }
        else if (quiet)
            launch_chained(h->stream, k_cell<false, true>, h->n_cell_blocks, kCellBlock, 0, a);
        else
            launch_chained(h->stream, k_cell<false, false>, h->n_cell_blocks, kCellBlock, 0, a);
    }
    h->shift_grid_pending = false;
    h->meas_src = nullptr;
    h->ranges_in_soa = false;
    float* t = h->free_cur;
    h->free_cur = h->free_next;
    h->free_next = t;
    return (int)cudaGetLastError();
}

int run_born_scan(dogm_handle* h)
{
    // (the dynamic-cell list is published by the birth kernel behind this scan; without birth particles, by the scan itself)
    return run_blocksum_scan(h, h->blk_sum, h->blk_off, h->n_cell_blocks, &h->scal->born_total, h->B <= 0);
}

int run_birth(dogm_handle* h)
{
    if (h->B > 0 && !h->band.enabled)
        return run_birth_fill(h, true); // the birth kernel scans the block sums itself
    int e = run_born_scan(h);
    return e ? e : run_birth_fill(h, false);
}

int run_birth_fill(dogm_handle* h, bool fused_scan)
{
    if (h->B <= 0) // (device-paced band cycle: the band's capacity, which sizes the grid)
        return 0;
    BirthArgs a;
    a.B = h->B;
    a.gs = h->gs;
    a.B_glob = h->band.enabled ? h->band.b_glob : h->B;
    a.cell_base = h->band.row0 * h->gs;
    a.slots = make_slot_view(h);
    a.scal = h->scal;
    a.born_masses = h->born_masses;
    a.meas = h->meas;
    a.grid = h->grid;
    a.birth = h->birth;
    a.noise = h->birth_noise;
    a.noise_injected = (h->opts.noise_mode == DOGM_NOISE_INJECTED) ? 1 : 0;
    a.stddev_velocity = h->params.stddev_velocity;
    a.seed = h->opts.seed + h->band.salt;
    a.cycle = h->cycle;
    a.pub_count = nullptr;
    a.pub_host = nullptr;
    a.pub_seq = 0;
    if (h->dyn_pub_armed && h->dyn_pub_dev)
    {
        a.pub_count = h->dyn_count;
        a.pub_host = h->dyn_pub_dev;
        a.pub_seq = ++h->dyn_pub_seq;
        h->dyn_pub_armed = false;
        h->dyn_pub_pending = true;
    }
    const int grid = div_up(h->B, kBlock);
    a.n_groups = 0;
    a.blk_off_w = h->blk_off;
    a.grp_word = h->grp_word;
    a.epoch = 0;
    a.total_out = &h->scal->born_total;
    a.cnt = h->band.dev_cnt;
    size_t smem = sizeof(double);
    if (fused_scan)
    {
        a.n_groups = h->n_blk_groups;
        a.epoch = ++h->birth_epoch;
        smem = ((size_t)h->n_blk_groups + 1) * sizeof(double);
    }
    {
        LaunchScope ls(h, K_BIRTH_PARTICLES, 25.0 * h->B + 16.0 * h->n_cell_blocks);
        launch_chained(h->stream, k_birth_particles, grid, kBlock, smem, a);
    }
    return (int)cudaGetLastError();
}

int run_extract_free_mass(dogm_handle* h)
{
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_extract_free, div_up(h->C, kBlock), kBlock, 0, h->grid, h->free_cur, h->C);
    return (int)cudaGetLastError();
}

int run_init_grid(dogm_handle* h)
{
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_init_grid, div_up(h->C, kBlock), kBlock, 0, h->grid, h->meas, h->free_cur, h->free_next,
                                                               h->cell_start, h->C);
    return (int)cudaGetLastError();
}

int run_extract_dynamic_cells(dogm_handle* h, float min_occ, float min_vel, dogm_dynamic_cell* d_out, int capacity,
                              int* d_count)
{
    LaunchScope ls(h, K_MISC, 0.0);
    cudaMemsetAsync(d_count, 0, sizeof(int), h->stream);
    launch_chained(h->stream, k_extract_dynamic, div_up(h->C, kBlock), kBlock, 0, h->grid, h->C, min_occ, min_vel, d_out, capacity,
                                                                     d_count);
    return (int)cudaGetLastError();
}

int trace_bind_cells(unsigned long long* p)
{
    return (int)cudaMemcpyToSymbol(c_trace, &p, sizeof(p));
}

} // namespace dogm_b200
