// Particle-side kernels of the DOGM cycle for sm_100a:
//   prediction (+ pass-0 digit histogram), stable LSD counting sort keyed on the bounded cell index,
//   per-cell segmented sums, persistent-weight update, joint-weight CDF, resampling search + gather.
//
// Reference behaviour restated here (paths relative to the reference's dogm/):
//   src/kernel/predict.cu:16-52, src/dogm.cu:262-281 (thrust::sort_by_key), src/kernel/particle_to_grid.cu:26-44,
//   src/kernel/update_persistent_particles.cu:49-87, src/dogm.cu:386-423, src/kernel/resampling.cu:17-68.
//
// Data layout inside a cycle: the public population lives in the reference's ParticlesSoA block (`pa`); prediction
// turns it into 32-byte records (PRec: one DRAM sector per particle) that stay in slot order; the sort passes move
// (cell, slot) pairs only, the segmented reduction and the resampling gather fetch the records through the sorted
// permutation; resampling writes the next population back into `pa`.
//
// All float arithmetic that feeds an index (cell index, birth slot, ancestor) is written with explicit
// round-to-nearest intrinsics and the file is compiled with -fmad=false, so the operation order below IS the
// numerical contract.
#include "dogm_internal.cuh"

namespace dogm_b200
{

// =========================================================================================================
// record helpers.  PRec halves: lo = (x, y, key, assoc), hi = (vx, vy, weight, pad): the segmented reduction only
// needs the hi half, the cell index travels with the position.
// =========================================================================================================
__device__ __forceinline__ float4 rec_lo(float x, float y, int key, uint32_t assoc)
{
    return make_float4(x, y, __int_as_float(key), __uint_as_float(assoc));
}
__device__ __forceinline__ float4 rec_hi(float vx, float vy, float w)
{
    return make_float4(vx, vy, w, 0.0f);
}

// =========================================================================================================
// prediction
// =========================================================================================================
struct PredictArgs
{
    const float4* state;
    const float* weight;
    const uint8_t* assoc;
    PRec* out;
    int* key_out;
    int n;
    int gs;            // row length of the (band of the) grid = cells per side of the whole grid
    int row0, rows;    // rows of the whole grid this handle owns (0, gs without bands)
    uint8_t* mig;      // band mode: per slot 0 = stays, 1 / 2 = leaves through the lower / upper edge (k_outbox_* compact them)
    float dt, p_S, sigma_pos, sigma_vel;
    int shift_active, x_move, y_move;
    const float4* noise;
    uint64_t seed;
    uint32_t cycle;
    uint32_t* hist; // [tiles][bins] pass-0 histogram
    int bins;
    uint32_t mask;
    double* leave_cnt[2]; // band mode: per tile, the particles leaving through the lower / upper edge (input of the outbox scan)
    uint32_t* clear[2];   // few tiles (fused sort passes): the tile histograms of the later passes, cleared here (filled with atomics)
    int clear_n[2];
    // bucket sort (BUCKET variant): bucket number per particle, tile histogram over the buckets
    uint16_t* bkt_out;  // [n] bucket number per slot
    const int* plan_key; // the cycle's splitters (BucketPlan, below)
    const int* plan_slot;
    const int* plan_base;
    const int* plan_pos;
    int plan_n, plan_shift;
};

// ---------------------------------------------------------------------------------------------------------
// Bucket sort, steps 1 and 2 (a one-CTA plan kernel, then inside the prediction kernel).  The population enters a cycle in cell
// order (resampling draws in the order of the sorted ancestors), and a prediction moves a particle by a few cells: the cell
// indices of the particles in slots 0, stride, 2 stride, ... are splitters that cut the NEW keys into buckets of about
// `stride` particles.  Where two splitters are more than kBucketSpan cells apart (sparse regions) the range is cut further, so
// that a bucket never spans more than kBucketSpan cells: the per-bucket counting sort (k_bucket_sort) then needs counters for
// kBucketSpan cells only.  Any splitters give a correct sort - they only decide how evenly the buckets fill.
// ---------------------------------------------------------------------------------------------------------
constexpr int kBucketSpan = 2048;
constexpr int kBucketSpanBits = 11;

// The splitters of a cycle (global memory, written by k_bucket_plan, read by k_predict through L1): sorted by (cell, slot), the
// order of the sort itself - splitting by slot as well keeps a cell with tens of thousands of particles from ending up in one
// bucket.  Range u holds the pairs from splitter u up to, not including, splitter u + 1 and is cut into nsub buckets of at
// most kBucketSpan cells.
struct BucketPlan
{
    int* key;   // [n_spl] cell of splitter u (key[0] = 0)
    int* slot;  // [n_spl] its slot (slot[0] = 0)
    int* base;  // [n_spl] number of the first bucket of range u
    int* pos;   // [n_spl] where the sample of slots stride * t .. ended up among the splitters (start of the search)
    int* org;   // [bins] first cell of every bucket
    int n_spl, stride_shift, bins;
};

// one value per thread of a 1024-thread CTA: inclusive scans
__device__ __forceinline__ int block1024_scan_max(int v, int* s_w)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d)
            v = max(v, t);
    }
    if (lane == 31)
        s_w[warp] = v;
    __syncthreads();
    int pre = INT_MIN;
    for (int w = 0; w < warp; w++)
        pre = max(pre, s_w[w]);
    __syncthreads();
    return max(v, pre);
}
__device__ __forceinline__ int block1024_scan_sum(int v, int* s_w, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d)
            v += t;
    }
    if (lane == 31)
        s_w[warp] = v;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < 32; w++)
    {
        if (w < warp)
            pre += s_w[w];
        tot += s_w[w];
    }
    __syncthreads();
    *total = tot;
    return v + pre;
}

// One CTA per cycle, in front of the prediction.  The samples are the cells (before this prediction, moved by the pending
// ego-motion shift) of the particles in slots 0, stride, 2 stride, ...  After a resampling the population is in cell order in
// two runs - the copies of persistent particles, then the copies of birth particles - so the samples are two sorted lists:
// they are merged by rank (one binary search per sample).  Any other input (a caller's particle set) gets a running maximum
// instead: still valid splitters, just less even buckets.
__global__ void __launch_bounds__(kWideBlock) k_bucket_plan(const int* __restrict__ smp_raw, BucketPlan pl, int delta, int C)
{
    pdl_prologue(K_MISC * 2 + 1);
    __shared__ int s_key[kWideBlock], s_slot[kWideBlock], s_raw[kWideBlock];
    __shared__ int s_w[32];
    __shared__ int s_descents, s_break;
    const int t = threadIdx.x, nt = pl.n_spl;
    if (t == 0)
    {
        s_descents = 0;
        s_break = nt;
    }
    int raw = INT_MAX;
    if (t < nt)
        raw = t == 0 ? 0 : min(max(__ldg(smp_raw + t) + delta, 0), C - 1);
    s_raw[t] = raw;
    __syncthreads();
    if (t > 0 && t < nt && raw < s_raw[t - 1])
    {
        atomicAdd(&s_descents, 1);
        atomicMin(&s_break, t);
    }
    __syncthreads();
    const int descents = s_descents, brk = s_break;
    int my_pos = t;
    if (descents == 1)
    { // two sorted runs [0, brk) and [brk, nt): rank of every sample in their merge (equal cells: the first run's come first)
        int lo, hi;
        if (t < brk)
        { // + samples of the second run with a smaller cell
            lo = brk, hi = nt;
            while (lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                if (s_raw[mid] < raw)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            my_pos = t + (lo - brk);
        }
        else if (t < nt)
        { // + samples of the first run with a cell <= mine
            lo = 0, hi = brk;
            while (lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                if (s_raw[mid] <= raw)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            my_pos = (t - brk) + lo;
        }
        if (t < nt)
        {
            s_key[my_pos] = raw;
            s_slot[my_pos] = t << pl.stride_shift;
        }
        __syncthreads();
    }
    else
    {
        const int mx = descents == 0 ? raw : block1024_scan_max(t < nt ? raw : INT_MIN, s_w);
        s_key[t] = t < nt ? mx : INT_MAX;
        s_slot[t] = t << pl.stride_shift;
        __syncthreads();
    }
    if (t < nt)
        pl.pos[t] = my_pos;
    // range t: from splitter t to splitter t + 1; the cell of splitter t + 1 itself can occur in it, hence the + 1
    const int lo_c = t < nt ? s_key[t] : 0;
    const int hi_c = t + 1 < nt ? s_key[t + 1] : C - 1;
    const int nsub = t < nt ? (hi_c - lo_c) / kBucketSpan + 1 : 0;
    int total;
    const int incl = block1024_scan_sum(nsub, s_w, &total);
    if (t < nt)
    {
        pl.key[t] = lo_c;
        pl.slot[t] = t == 0 ? 0 : s_slot[t];
        pl.base[t] = incl - nsub;
        for (int k = 0; k < nsub; k++)
            if (incl - nsub + k < pl.bins)
                pl.org[incl - nsub + k] = lo_c + k * kBucketSpan;
    }
}

// splitter u <= (key, slot) in the order of the sort?
__device__ __forceinline__ bool splitter_le(const BucketPlan& pl, int u, int key, int slot)
{
    const int c = __ldg(pl.key + u);
    return c < key || (c == key && __ldg(pl.slot + u) <= slot);
}

// the bucket of (key, slot): the last range whose splitter is <= the pair (splitter 0 = (0, 0) is <= everything).  A prediction
// moves a particle by a few cells, so the search gallops away from the splitter its own slot's sample became: two or three
// probes as a rule, all of them L1 hits (the particles of a CTA search the same neighbourhood).
__device__ __forceinline__ int bucket_of(const BucketPlan& pl, int key, int slot)
{
    const int nt = pl.n_spl;
    int lo, hi; // invariant: splitter lo <= item < splitter hi (hi == nt: none)
    const int u0 = __ldg(pl.pos + min(slot >> pl.stride_shift, nt - 1));
    if (splitter_le(pl, u0, key, slot))
    {
        lo = u0;
        int step = 1;
        hi = nt;
        while (lo + step < nt)
        {
            if (!splitter_le(pl, lo + step, key, slot))
            {
                hi = lo + step;
                break;
            }
            lo += step;
            step <<= 1;
        }
    }
    else
    {
        hi = u0;
        int step = 1;
        lo = 0;
        while (hi - step > 0)
        {
            if (splitter_le(pl, hi - step, key, slot))
            {
                lo = hi - step;
                break;
            }
            hi -= step;
            step <<= 1;
        }
    }
    while (hi - lo > 1)
    {
        const int mid = (lo + hi) >> 1;
        if (splitter_le(pl, mid, key, slot))
            lo = mid;
        else
            hi = mid;
    }
    return __ldg(pl.base + lo) + ((key - __ldg(pl.key + lo)) >> kBucketSpanBits);
}

__device__ __forceinline__ float4 predict_noise_philox(uint64_t seed, uint32_t slot, uint32_t cycle, float sigma_pos,
                                                       float sigma_vel)
{
    const float4 g = philox_normal4(seed, slot, STAGE_PREDICT, cycle);
    return make_float4(__fmul_rn(g.x, sigma_pos), __fmul_rn(g.y, sigma_pos), __fmul_rn(g.z, sigma_vel),
                       __fmul_rn(g.w, sigma_vel));
}

// predictKernel, predict.cu:16-52.  The ego-motion particle shift of moveParticlesKernel
// (ego_motion_compensation.cu:16-23) is applied first when a shift is pending: same two roundings as the
// reference's separate kernel.  x' = (x + dt*vx) + noise: three roundings, the GLM mat4*vec4 grouping of predict.cu:33.
// One CTA of 1024 threads per sort tile (4 particles per thread): the tile's pass-0 digit histogram comes for free.
// Output: the particle as a 32-byte record in slot order (it never moves again inside the cycle) and its cell
// index in a compact key array (the only thing the sort passes touch).  BUCKET: also the particle's bucket number, and the
// tile histogram is over the buckets (bucket sort, above).
template <bool INJECTED, bool BUCKET>
__global__ void __launch_bounds__(kWideBlock) k_predict(PredictArgs a)
{
    pdl_prologue(K_PREDICT * 2);
    extern __shared__ uint32_t s_hist[];
    __shared__ uint32_t s_leave[2];
    constexpr int kPer = kTileItems / kWideBlock;
    BucketPlan pl;
    pl.key = const_cast<int*>(a.plan_key);
    pl.slot = const_cast<int*>(a.plan_slot);
    pl.base = const_cast<int*>(a.plan_base);
    pl.pos = const_cast<int*>(a.plan_pos);
    pl.org = nullptr;
    pl.n_spl = a.plan_n;
    pl.stride_shift = a.plan_shift;
    pl.bins = a.bins;

    const float4* __restrict__ state = a.state;
    const float* __restrict__ weight = a.weight;
    const uint8_t* __restrict__ assoc = a.assoc;
    const float4* __restrict__ noise = a.noise;
    PRec* __restrict__ out = a.out;
    int* __restrict__ key_out = a.key_out;
    const int base = blockIdx.x * kTileItems;
    const float hi = (float)(a.gs - 1);
    const float xm = (float)a.x_move, ym = (float)a.y_move;
    const uint64_t keep = l2_evict_last(); // the records are gathered twice later in the cycle

    for (int b = threadIdx.x; b < a.bins; b += kWideBlock)
        s_hist[b] = 0u;
    if (threadIdx.x < 2)
        s_leave[threadIdx.x] = 0u;
#pragma unroll
    for (int k = 0; k < 2; k++)
        for (int b = blockIdx.x * kWideBlock + threadIdx.x; b < a.clear_n[k]; b += gridDim.x * kWideBlock)
            a.clear[k][b] = 0u;
    __syncthreads();

#pragma unroll
    for (int j = 0; j < kPer; j++)
    {
        const int i = base + j * kWideBlock + threadIdx.x;
        if (i < a.n)
        {
            float4 s = state[i];
            float w = weight[i];
            const uint32_t as = assoc[i];
            float4 nz;
            if (INJECTED)
                nz = noise[i];
            else
                nz = predict_noise_philox(a.seed, (uint32_t)i, a.cycle, a.sigma_pos, a.sigma_vel);
            if (a.shift_active)
            {
                s.x = __fsub_rn(s.x, xm);
                s.y = __fsub_rn(s.y, ym);
            }
            const float x = __fadd_rn(__fadd_rn(s.x, __fmul_rn(a.dt, s.z)), nz.x);
            const float y = __fadd_rn(__fadd_rn(s.y, __fmul_rn(a.dt, s.w)), nz.y);
            const float vx = __fadd_rn(s.z, nz.z);
            const float vy = __fadd_rn(s.w, nz.w);
            w = __fmul_rn(a.p_S, w);
            if ((x > hi || x < 0.0f) || (y > hi || y < 0.0f))
                w = 0.0f;
            const int px = min(max(__float2int_rz(x), 0), a.gs - 1);
            int py = min(max(__float2int_rz(y), 0), a.gs - 1) - a.row0;
            float w_leaving = 0.0f;
            uint8_t leaves = 0;
            if (py < 0 || py >= a.rows)
            { // band mode only: the particle now lies in a neighbour's band.  It stays behind as a weightless ghost in the
              // edge row until resampling drops it; its weight rides in the record's spare word and the slot is flagged, so
              // that k_outbox_write can copy it into the outbox IN SLOT ORDER (the order of an atomic append would change
              // from run to run, and with it the neighbour's particle order)
                if (w > 0.0f)
                {
                    leaves = py < 0 ? 1 : 2;
                    w_leaving = w;
                }
                w = 0.0f;
                py = py < 0 ? 0 : a.rows - 1;
            }
            if (a.mig)
            {
                a.mig[i] = leaves;
                if (leaves)
                    atomicAdd(&s_leave[leaves - 1], 1u); // (few particles cross an edge: a handful of atomics per tile)
            }
            const int cell = px + a.gs * py;
            float4* o = reinterpret_cast<float4*>(out + i);
#ifdef DOGM_NO_L2_HINTS
            o[0] = rec_lo(x, y, cell, as);
            o[1] = make_float4(vx, vy, w, w_leaving);
#else
            st_hint(o, rec_lo(x, y, cell, as), keep);
            st_hint(o + 1, make_float4(vx, vy, w, w_leaving), keep);
#endif
            key_out[i] = cell;
            if (BUCKET)
            {
                const int bk = bucket_of(pl, cell, i);
                a.bkt_out[i] = (uint16_t)bk;
                atomicAdd(&s_hist[bk], 1u);
            }
            else
                atomicAdd(&s_hist[(uint32_t)cell & a.mask], 1u);
        }
    }
    __syncthreads();
    uint32_t* row = a.hist + (size_t)blockIdx.x * a.bins;
    for (int b = threadIdx.x; b < a.bins; b += kWideBlock)
        row[b] = s_hist[b];
    if (a.mig && threadIdx.x < 2)
        a.leave_cnt[threadIdx.x][blockIdx.x] = (double)s_leave[threadIdx.x];
}

// splitter samples straight from the particle block (first cycle, after dogm_set_particles: resampling has not left them)
__global__ void __launch_bounds__(kBlock) k_sample_keys(const int* __restrict__ idx, int n, int* smp_raw, int stride_shift)
{
    pdl_prologue(K_MISC * 2);
    const int u = blockIdx.x * kBlock + threadIdx.x;
    if (((long long)u << stride_shift) < n)
        smp_raw[u] = idx[(size_t)u << stride_shift];
}

// SoA -> records without prediction (+ pass-0 histogram): the entry into the sort when the keys did not come from
// k_predict, e.g. dogm_particle_assignment right after dogm_set_particles
__global__ void __launch_bounds__(kWideBlock) k_soa_to_rec(const float4* __restrict__ state, const int* __restrict__ idx,
                                                           const float* __restrict__ weight,
                                                           const uint8_t* __restrict__ assoc, PRec* __restrict__ out,
                                                           int* __restrict__ key_out, int n, uint32_t* hist, int bins,
                                                           uint32_t mask)
{
    pdl_prologue(K_TILE_HIST * 2);
    extern __shared__ uint32_t s_hist[];
    for (int b = threadIdx.x; b < bins; b += kWideBlock)
        s_hist[b] = 0u;
    __syncthreads();
    const int base = blockIdx.x * kTileItems;
#pragma unroll
    for (int j = 0; j < kTileItems / kWideBlock; j++)
    {
        const int i = base + j * kWideBlock + threadIdx.x;
        if (i < n)
        {
            const int key = idx[i];
            const float4 s = state[i];
            float4* o = reinterpret_cast<float4*>(out + i);
            o[0] = rec_lo(s.x, s.y, key, assoc[i]);
            o[1] = rec_hi(s.z, s.w, weight[i]);
            key_out[i] = key;
            atomicAdd(&s_hist[(uint32_t)key & mask], 1u);
        }
    }
    __syncthreads();
    uint32_t* row = hist + (size_t)blockIdx.x * bins;
    for (int b = threadIdx.x; b < bins; b += kWideBlock)
        row[b] = s_hist[b];
}

// pass-0 histogram of keys that are already in place (assignment called twice in a row)
__global__ void __launch_bounds__(kWideBlock) k_key_tile_hist(const int* __restrict__ key, int n, uint32_t* hist, int bins,
                                                              uint32_t mask, const int* n_dev)
{
    pdl_prologue(K_TILE_HIST * 2);
    if (n_dev)
        n = *n_dev; // device-paced band cycle: the count is only known on the device; the grid is an estimate, the CTAs loop
    extern __shared__ uint32_t s_hist[];
    const int tiles = (n + kTileItems - 1) / kTileItems;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        for (int b = threadIdx.x; b < bins; b += kWideBlock)
            s_hist[b] = 0u;
        __syncthreads();
        const int base = tile * kTileItems;
#pragma unroll
        for (int j = 0; j < kTileItems / kWideBlock; j++)
        {
            const int i = base + j * kWideBlock + threadIdx.x;
            if (i < n)
                atomicAdd(&s_hist[(uint32_t)key[i] & mask], 1u);
        }
        __syncthreads();
        uint32_t* row = hist + (size_t)tile * bins;
        for (int b = threadIdx.x; b < bins; b += kWideBlock)
            row[b] = s_hist[b];
        __syncthreads();
    }
}

// histogram of a later pass over the pairs the previous pass has just written (they are still in L2); cheaper than
// 2e6 global atomics issued from the previous scatter (21 us against 6 us at 2e6 particles)
__global__ void __launch_bounds__(kWideBlock) k_pair_tile_hist(const int2* __restrict__ pair, int n, uint32_t* hist, int bins,
                                                               int shift, uint32_t mask, const int* n_dev)
{
    pdl_prologue(K_TILE_HIST * 2 + 1);
    if (n_dev)
        n = *n_dev;
    extern __shared__ uint32_t s_hist[];
    const int tiles = (n + kTileItems - 1) / kTileItems;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        for (int b = threadIdx.x; b < bins; b += kWideBlock)
            s_hist[b] = 0u;
        __syncthreads();
        const int base = tile * kTileItems;
        int keys[kTileItems / kWideBlock];
#pragma unroll
        for (int j = 0; j < kTileItems / kWideBlock; j++)
        {
            const int i = base + j * kWideBlock + threadIdx.x;
            keys[j] = i < n ? __ldg(&pair[i].x) : -1;
        }
#pragma unroll
        for (int j = 0; j < kTileItems / kWideBlock; j++)
            if (keys[j] >= 0)
                atomicAdd(&s_hist[((uint32_t)keys[j] >> shift) & mask], 1u);
        __syncthreads();
        uint32_t* row = hist + (size_t)tile * bins;
        for (int b = threadIdx.x; b < bins; b += kWideBlock)
            row[b] = s_hist[b];
        __syncthreads();
    }
}

// records -> the reference's SoA block (read-out between stages: getParticles after prediction / assignment);
// with `spair` the output is in sorted order (position p holds record spair[p].y)
__global__ void __launch_bounds__(kBlock) k_rec_to_soa(const PRec* __restrict__ rec, const int2* __restrict__ spair,
                                                       float4* __restrict__ state, int* __restrict__ idx,
                                                       float* __restrict__ weight, uint8_t* __restrict__ assoc, int n)
{
    pdl_prologue(K_MISC * 2);
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n)
        return;
    const int slot = spair ? spair[i].y : i;
    const float4* r = reinterpret_cast<const float4*>(rec + slot);
    const float4 lo = r[0], hi = r[1];
    state[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
    idx[i] = __float_as_int(lo.z);
    weight[i] = hi.z;
    assoc[i] = (uint8_t)__float_as_uint(lo.w);
}

// =========================================================================================================
// counting sort: scan of the [tiles][bins] digit histograms
//   in : table[t][b] = number of keys of tile t whose digit is b
//   out: table[t][b] = number of keys with digit b in tiles < t;   bin_base[b] = number of keys with digit < b
// One CTA of 32 warps per 32 bins (lane = bin, coalesced 128-byte rows); warp w owns a contiguous range of tiles
// and keeps 16 independent loads in flight.  The last CTA to finish scans the bin totals.  As a side job every
// CTA clears its share of the NEXT pass's table (filled with atomics by the scatter kernel that follows).
// =========================================================================================================
constexpr int kScanWarps = kWideBlock / 32;
constexpr int kScanBatch = 16;

// integer flavour of publish_f64 / await_f64 (values below 2^63; the top bit carries the launch parity)
__device__ __forceinline__ void publish_u64(unsigned long long* p, unsigned long long v, uint32_t epoch)
{
    const unsigned long long bits = (v & 0x7fffffffffffffffull) | ((unsigned long long)(epoch & 1u) << 63);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(bits) : "memory");
}
__device__ __forceinline__ unsigned long long await_u64(const unsigned long long* p, uint32_t epoch)
{
    unsigned long long bits;
    for (;;)
    {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(bits) : "l"(p) : "memory");
        if ((uint32_t)(bits >> 63) == (epoch & 1u))
            break;
        __nanosleep(40);
    }
    return bits & 0x7fffffffffffffffull;
}

// CTA j owns the 32 bins [32 j, 32 j + 32): lane = bin, the warps split the tiles.  On exit table[t][b] is the final
// position of the first key of tile t with digit b: (keys with a smaller digit) + (same digit in earlier tiles).  The
// first term needs the totals of the CTAs before: every CTA publishes the sum of its 32 bins and adds up the words of
// its predecessors (at most 64 CTAs, all resident; integers, so the order of the additions does not matter).
__global__ void __launch_bounds__(kWideBlock) k_hist_scan(uint32_t* __restrict__ table, int tiles, int bins,
                                                          unsigned long long* chain, uint32_t epoch, const int* n_dev)
{
    pdl_prologue(K_HIST_SCAN * 2 + (bins > 1024 ? 0 : 1));
    if (n_dev)
        tiles = (*n_dev + kTileItems - 1) / kTileItems;
    __shared__ uint32_t s_part[kScanWarps][32];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bin = blockIdx.x * 32 + lane;
    const int per = (tiles + kScanWarps - 1) / kScanWarps;
    const int t0 = min(warp * per, tiles), t1 = min(t0 + per, tiles);

    // the common case (at most 16 tiles per warp, i.e. up to 2e6 particles): one batch of loads, kept in registers
    uint32_t v[kScanBatch];
    const bool single = (t1 - t0) <= kScanBatch;
    uint32_t sum = 0;
    if (single)
    {
#pragma unroll
        for (int k = 0; k < kScanBatch; k++)
            v[k] = (t0 + k < t1) ? table[(size_t)(t0 + k) * bins + bin] : 0u;
#pragma unroll
        for (int k = 0; k < kScanBatch; k++)
            sum += v[k];
    }
    else
    {
        for (int tb = t0; tb < t1; tb += kScanBatch)
        {
#pragma unroll
            for (int k = 0; k < kScanBatch; k++)
                v[k] = (tb + k < t1) ? table[(size_t)(tb + k) * bins + bin] : 0u;
#pragma unroll
            for (int k = 0; k < kScanBatch; k++)
                sum += v[k];
        }
    }
    s_part[warp][lane] = sum;
    __syncthreads();
    uint32_t run = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; w++)
    {
        const uint32_t p = s_part[w][lane];
        if (w < warp)
            run += p;
        total += p;
    }
    // exclusive scan of the 32 bin totals inside the warp, CTA total to the chain
    uint32_t incl = total;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += t;
    }
    if (threadIdx.x == 31)
        publish_u64(chain + blockIdx.x, incl, epoch);
    if (warp == 0)
    {
        uint32_t before = 0;
        for (int j = lane; j < (int)blockIdx.x; j += 32)
            before += (uint32_t)await_u64(chain + j, epoch);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            before += __shfl_xor_sync(0xffffffffu, before, d);
        if (lane == 0)
            s_base = before;
    }
    __syncthreads();
    run += s_base + (incl - total);
    if (single)
    {
#pragma unroll
        for (int k = 0; k < kScanBatch; k++)
            if (t0 + k < t1)
            {
                table[(size_t)(t0 + k) * bins + bin] = run;
                run += v[k];
            }
    }
    else
    {
        for (int tb = t0; tb < t1; tb += kScanBatch)
        {
#pragma unroll
            for (int k = 0; k < kScanBatch; k++)
                v[k] = (tb + k < t1) ? table[(size_t)(tb + k) * bins + bin] : 0u;
#pragma unroll
            for (int k = 0; k < kScanBatch; k++)
                if (tb + k < t1)
                {
                    table[(size_t)(tb + k) * bins + bin] = run;
                    run += v[k];
                }
        }
    }
}

// =========================================================================================================
// counting sort: stable scatter of one digit pass over (key, slot) pairs (one CTA per tile of 4096 items).
// Only these 8-byte pairs move; the 32-byte particle records stay where prediction wrote them.
// rank of an item = bin_base[digit] + (same digit in earlier tiles) + (same digit in earlier warps of the tile)
//                   + (same digit earlier in its own warp), the last term by __match_any ranking.
// =========================================================================================================
struct ScatterArgs
{
    const int* key_in;   // first pass: keys in slot order (slot = index)
    const int2* pair_in; // later passes: (key, slot)
    int2* pair_out;
    int n;
    int shift;
    uint32_t mask;
    int bins;
    const uint32_t* table; // this pass: final position of the first key per (tile, digit), from k_hist_scan
    const int* n_dev;      // device-paced band cycle: the item count lives on the device (n is an upper bound for the grid)
    const uint16_t* digit_in; // SRC 2: the digit (bucket number) per slot
    // FUSE (few tiles: the cycle is bound by launch latency): `table` holds the raw counts [tiles][bins] and every CTA scans the
    // columns itself; the tile histogram of the NEXT pass is added up with atomics while the pairs are stored (next_hist was
    // cleared by the prediction kernel), so neither k_hist_scan nor k_pair_tile_hist is launched
    int tiles;
    uint32_t* next_hist; // [tiles][next_bins] or null (last pass)
    int next_bins, next_shift;
};

// SRC 0: first radix pass (keys in slot order, digits nearly all different: ballot ranking)
// SRC 1: later radix passes ((key, slot) pairs grouped by cell already: MATCH.ANY ranking)
// SRC 2: grouping pass of the bucket sort (keys in slot order + the bucket number k_predict computed per slot; a warp's
//        particles fall into a handful of buckets: MATCH.ANY ranking, long contiguous runs in the output)
template <int SRC, bool FUSE>
__device__ __forceinline__ void scatter_tile(const ScatterArgs& a, const int tile, unsigned short* s_cnt /* [warps][bins] */)
{
    constexpr bool FIRST = SRC == 0;
    constexpr bool DIG = SRC == 2;
    const uint16_t* __restrict__ digit_in = a.digit_in;

    const int* __restrict__ key_in = a.key_in;
    const int2* __restrict__ pair_in = a.pair_in;
    int2* __restrict__ pair_out = a.pair_out;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    const int warp_base = tile * kTileItems + warp * (kTileItems / kWarpsPerBlock);

    // all key loads of the warp's 16 rounds are in flight while the rank counters are cleared
    int keys[kRoundsPerWarp];
#pragma unroll
    for (int r = 0; r < kRoundsPerWarp; r++)
    {
        const int i = warp_base + r * 32 + lane;
        keys[r] = (i < a.n) ? (DIG ? (int)digit_in[i] : (FIRST ? key_in[i] : pair_in[i].x)) : 0; // (DIG: shift 0, mask 0xffff)
    }
    {
        uint4* z = reinterpret_cast<uint4*>(s_cnt); // bins is a multiple of 32: the counter block is a multiple of 16 B
        const int nz = a.bins * kWarpsPerBlock * (int)sizeof(unsigned short) / (int)sizeof(uint4);
        for (int b = threadIdx.x; b < nz; b += kBlock)
            z[b] = make_uint4(0u, 0u, 0u, 0u);
    }
    // FUSE: final position of this tile's first key per digit = (keys with a smaller digit) + (same digit in earlier tiles),
    // from the raw counts of all tiles (a few dozen rows, read by every CTA: they sit in L2)
    uint32_t* s_row = reinterpret_cast<uint32_t*>(s_cnt + a.bins * kWarpsPerBlock); // [bins] (FUSE only)
    if (FUSE)
    {
        uint32_t* s_tot = s_row + a.bins;
        __shared__ uint32_t s_ws[kWarpsPerBlock];
        for (int b = threadIdx.x; b < a.bins; b += kBlock)
        {
            uint32_t before = 0, total = 0;
            const uint32_t* col = a.table + b;
            int t = 0;
            for (; t + 16 <= a.tiles; t += 16)
            { // (16 rows in flight per thread: the loop is a chain of L2 round trips)
                uint32_t v[16];
#pragma unroll
                for (int k = 0; k < 16; k++)
                    v[k] = __ldcg(col + (size_t)(t + k) * a.bins);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    total += v[k];
                    before += t + k < tile ? v[k] : 0u;
                }
            }
            for (; t < a.tiles; t++)
            {
                const uint32_t v = __ldcg(col + (size_t)t * a.bins);
                total += v;
                before += t < tile ? v : 0u;
            }
            s_row[b] = before;
            s_tot[b] = total;
        }
        __syncthreads();
        // exclusive scan of the digit totals: thread t owns `per` consecutive digits (bins <= 4 * 256)
        const int per = a.bins >= kBlock ? a.bins / kBlock : 1;
        const int e0 = (int)threadIdx.x * per;
        uint32_t cnt[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            cnt[k] = (k < per && e0 + k < a.bins) ? s_tot[e0 + k] : 0u;
            sum += cnt[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t u = __shfl_up_sync(full, incl, d);
            if (lane >= d)
                incl += u;
        }
        if (lane == 31)
            s_ws[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int w = 0; w < warp; w++)
            run += s_ws[w];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < per && e0 + k < a.bins)
            {
                s_row[e0 + k] += run;
                run += cnt[k];
            }
    }
    __syncthreads();

    unsigned short* my_cnt = s_cnt + warp * a.bins;
    uint32_t rank2[kRoundsPerWarp / 2]; // ranks inside the warp's 512 items, two per register

    // lanes with the same digit, for all 16 rounds first: the MATCH results are independent of each other and of the
    // counters, so their latencies overlap instead of adding up round by round
    uint32_t peers[kRoundsPerWarp];
#pragma unroll
    for (int r = 0; r < kRoundsPerWarp; r++)
    {
        const bool valid = warp_base + r * 32 + lane < a.n;
        const uint32_t digit = valid ? (((uint32_t)keys[r] >> a.shift) & a.mask) : 0xffffffffu;
        if (!FIRST)
        {
            peers[r] = __match_any_sync(full, digit);
        }
        else
        {
            // MATCH.ANY takes time in proportion to the number of distinct digits in the warp: close to 32 in the first
            // pass, where neighbouring particles have just been scattered by the process noise (31 us against 22 us
            // at 2e6 particles), few in the later passes, whose input is grouped by cell already (10 us against 19 us).
            // One ballot per digit bit costs the same for any input.
            uint32_t p = __ballot_sync(full, valid);
#pragma unroll
            for (int b = 0; b < kMaxDigitBits; b++)
            {
                const bool bit = (digit >> b) & 1u;
                const uint32_t m = __ballot_sync(full, bit);
                p &= bit ? m : ~m;
            }
            peers[r] = p;
        }
    }
#pragma unroll
    for (int r = 0; r < kRoundsPerWarp; r++)
    {
        const bool valid = warp_base + r * 32 + lane < a.n;
        const uint32_t digit = ((uint32_t)keys[r] >> a.shift) & a.mask;
        // every lane reads its digit's counter, the first lane of each group then advances it
        uint32_t old = 0;
        if (valid)
            old = my_cnt[digit];
        __syncwarp(); // all reads of this round before the group leaders' writes
        if (valid && (peers[r] & lt) == 0)
            my_cnt[digit] = (unsigned short)(old + __popc(peers[r]));
        const uint32_t rk = old + __popc(peers[r] & lt);
        if (r & 1)
            rank2[r >> 1] |= rk << 16;
        else
            rank2[r >> 1] = rk;
        __syncwarp();
    }
    __syncthreads();

    // per bin: exclusive prefix over the warps of the tile; thread t owns the bins t, t + 256, ... (conflict-free)
    for (int b = threadIdx.x; b < a.bins; b += kBlock)
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; w++)
        {
            const uint32_t c = s_cnt[w * a.bins + b];
            s_cnt[w * a.bins + b] = (unsigned short)run;
            run += c;
        }
    }
    __syncthreads();

    // store phase: the inputs are read again (L1 / L2 hits) instead of being kept in registers across the ranking;
    // all loads of a batch of 8 are issued before the first store (the stores could alias them for the compiler)
    const uint32_t* __restrict__ row = FUSE ? s_row : a.table + (size_t)tile * a.bins; // final position of this tile's first key per bin
#pragma unroll
    for (int half = 0; half < 2; half++)
    {
        int key[8], slot[8];
        uint32_t dig[DIG ? 8 : 1];
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            const int i = warp_base + (half * 8 + q) * 32 + lane;
            key[q] = 0;
            slot[q] = i;
            if (DIG)
                dig[q] = 0;
            if (i < a.n)
            {
                if (FIRST || DIG)
                {
                    key[q] = __ldg(key_in + i);
                    if (DIG)
                        dig[q] = __ldg(digit_in + i);
                }
                else
                {
                    const int2 p = __ldg(pair_in + i);
                    key[q] = p.x;
                    slot[q] = p.y;
                }
            }
        }
        uint32_t dest[8];
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            const int r = half * 8 + q;
            const uint32_t digit = DIG ? dig[DIG ? q : 0] : (((uint32_t)key[q] >> a.shift) & a.mask);
            const uint32_t rk = (r & 1) ? (rank2[r >> 1] >> 16) : (rank2[r >> 1] & 0xffffu);
            dest[q] = (FUSE ? row[digit] : __ldg(row + digit)) + my_cnt[digit] + rk;
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            const int i = warp_base + (half * 8 + q) * 32 + lane;
            if (i < a.n)
            {
                pair_out[dest[q]] = make_int2(key[q], slot[q]);
                if (FUSE && a.next_hist)
                    atomicAdd(a.next_hist + (size_t)(dest[q] / kTileItems) * a.next_bins +
                                  (((uint32_t)key[q] >> a.next_shift) & (uint32_t)(a.next_bins - 1)),
                              1u);
            }
        }
    }
}

// LOOP = false: one CTA per tile (the grid is the number of tiles).  LOOP = true (device-paced band cycle): the item count is
// read from the device, the grid is the host's estimate and the CTAs walk over the tiles.
template <int SRC, bool LOOP, bool FUSE = false>
__global__ void __launch_bounds__(kBlock, 4) k_scatter(const ScatterArgs a_in)
{
    pdl_prologue(K_SCATTER * 2 + (SRC == 1 ? 1 : 0));
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned short* s_cnt = (unsigned short*)s_raw;
    if (!LOOP)
    {
        scatter_tile<SRC, FUSE>(a_in, (int)blockIdx.x, s_cnt);
        return;
    }
    ScatterArgs a = a_in;
    if (a.n_dev)
        a.n = *a.n_dev;
    const int tiles = (a.n + kTileItems - 1) / kTileItems;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        scatter_tile<SRC, false>(a, tile, s_cnt);
        __syncthreads();
    }
}

// =========================================================================================================
// Bucket sort, step 3: one CTA per bucket sorts its (key, slot) pairs by key, keeping the slot order inside a key (the pairs
// of a bucket arrive in slot order from the grouping pass).  A bucket spans at most kBucketSpan cells, so this is one counting
// pass over the local key  key - first cell of the bucket:  rank inside the warp by ballots (the local keys of neighbouring
// slots are nearly all different), per-warp u16 counters, prefix over the warps, exclusive scan over the cells.  The usual
// bucket (up to 4096 pairs) is sorted into shared memory and written out as one contiguous stretch; a larger one (a cell
// with thousands of particles) goes chunk by chunk straight to its final positions.
// Replaces the second radix pass with its tile histogram and table scan; no table, no global counters.
// =========================================================================================================
struct BucketSortArgs
{
    const int2* pair_in;    // grouped by bucket, slot order inside a bucket
    int2* pair_out;         // sorted by (cell, slot)
    const uint32_t* start;  // [bins] first position of every bucket (row 0 of the scanned grouping table)
    const int* org;         // [bins] first cell of every bucket
    int bins, n;
};

constexpr int kBucketSmemBytes = kWarpsPerBlock * kBucketSpan * 2 + 2 * kBucketSpan * 4 + kTileItems * 2 + kTileItems * 4;

// ranks of one chunk of up to 4096 pairs starting at `in` (the warp's 512 in 16 rounds).  On exit: rank2 = rank of every item among
// the items of its warp with the same local key (two per register), s_cnt[w][k] = items with local key k in the warps before w,
// s_tot[k] = items of the chunk with local key k.
__device__ __forceinline__ void bucket_rank_chunk(const int2* __restrict__ in, const int cnt, const int org, unsigned short* s_cnt,
                                                  uint32_t* s_tot, uint32_t (&rank2)[kRoundsPerWarp / 2])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu, lt = lanemask_lt();
    const int wb = warp * (kTileItems / kWarpsPerBlock);
    int lk[kRoundsPerWarp];
#pragma unroll
    for (int r = 0; r < kRoundsPerWarp; r++)
    {
        const int j = wb + r * 32 + lane;
        lk[r] = j < cnt ? __ldg(&in[j].x) - org : 0;
    }
    {
        uint4* z = reinterpret_cast<uint4*>(s_cnt);
        for (int b = threadIdx.x; b < kWarpsPerBlock * kBucketSpan * 2 / 16; b += kBlock)
            z[b] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    unsigned short* my_cnt = s_cnt + warp * kBucketSpan;
    uint32_t peers[kRoundsPerWarp];
#pragma unroll
    for (int r = 0; r < kRoundsPerWarp; r++)
    {
        const bool valid = wb + r * 32 + lane < cnt;
        uint32_t p = __ballot_sync(full, valid);
#pragma unroll
        for (int b = 0; b < kBucketSpanBits; b++)
        {
            const bool bit = ((uint32_t)lk[r] >> b) & 1u;
            const uint32_t m = __ballot_sync(full, bit);
            p &= bit ? m : ~m;
        }
        peers[r] = p;
    }
#pragma unroll
    for (int r = 0; r < kRoundsPerWarp; r++)
    {
        const bool valid = wb + r * 32 + lane < cnt;
        uint32_t old = 0;
        if (valid)
            old = my_cnt[lk[r]];
        __syncwarp();
        if (valid && (peers[r] & lt) == 0)
            my_cnt[lk[r]] = (unsigned short)(old + __popc(peers[r]));
        const uint32_t rk = old + __popc(peers[r] & lt);
        if (r & 1)
            rank2[r >> 1] |= rk << 16;
        else
            rank2[r >> 1] = rk;
        __syncwarp();
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kBucketSpan; b += kBlock)
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; w++)
        {
            const uint32_t c = s_cnt[w * kBucketSpan + b];
            s_cnt[w * kBucketSpan + b] = (unsigned short)run;
            run += c;
        }
        s_tot[b] = run;
    }
    __syncthreads();
}

// exclusive scan of the kBucketSpan counters in place (thread t owns the 8 consecutive counters 8 t .. 8 t + 7)
__device__ __forceinline__ void bucket_scan_counts(uint32_t* s_tot, uint32_t* s_ws)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* v = reinterpret_cast<uint4*>(s_tot) + 2 * threadIdx.x;
    uint4 a = v[0], b = v[1];
    const uint32_t sum = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += t;
    }
    if (lane == 31)
        s_ws[warp] = incl;
    __syncthreads();
    uint32_t pre = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; w++)
        if (w < warp)
            pre += s_ws[w];
    uint32_t run = pre + incl - sum;
    uint4 oa, ob;
    oa.x = run; run += a.x;
    oa.y = run; run += a.y;
    oa.z = run; run += a.z;
    oa.w = run; run += a.w;
    ob.x = run; run += b.x;
    ob.y = run; run += b.y;
    ob.z = run; run += b.z;
    ob.w = run;
    v[0] = oa;
    v[1] = ob;
    __syncthreads();
}

__global__ void __launch_bounds__(kBlock, 3) k_bucket_sort(const BucketSortArgs a)
{
    pdl_prologue(K_SCATTER * 2 + 1);
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned short* s_cnt = (unsigned short*)s_raw;                                   // [warps][span]
    uint32_t* s_tot = (uint32_t*)(s_raw + kWarpsPerBlock * kBucketSpan * 2);          // [span] counts, then offsets
    uint32_t* s_fill = s_tot + kBucketSpan;                                           // [span] (large buckets)
    uint32_t* s_sl = s_fill + kBucketSpan;                                            // [4096] sorted slots
    unsigned short* s_lk = (unsigned short*)(s_sl + kTileItems);                      // [4096] sorted local keys
    __shared__ uint32_t s_ws[kWarpsPerBlock];
    const int b = blockIdx.x;
    const int p0 = (int)a.start[b];
    const int p1 = b + 1 < a.bins ? (int)a.start[b + 1] : a.n;
    const int count = p1 - p0;
    if (count <= 0)
        return;
    const int org = a.org[b];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wb = warp * (kTileItems / kWarpsPerBlock);
    const int2* __restrict__ in = a.pair_in + p0;
    uint32_t rank2[kRoundsPerWarp / 2];
    if (count <= kTileItems)
    {
        bucket_rank_chunk(in, count, org, s_cnt, s_tot, rank2);
        bucket_scan_counts(s_tot, s_ws);
        const unsigned short* my_cnt = s_cnt + warp * kBucketSpan;
#pragma unroll
        for (int half = 0; half < 2; half++)
        {
            int2 pr[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                const int j = wb + (half * 8 + q) * 32 + lane;
                pr[q] = j < count ? __ldg(in + j) : make_int2(org, 0);
            }
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                const int r = half * 8 + q;
                const int j = wb + r * 32 + lane;
                if (j < count)
                {
                    const int k = pr[q].x - org;
                    const uint32_t rk = (r & 1) ? (rank2[r >> 1] >> 16) : (rank2[r >> 1] & 0xffffu);
                    const uint32_t lp = s_tot[k] + my_cnt[k] + rk;
                    s_sl[lp] = (uint32_t)pr[q].y;
                    s_lk[lp] = (unsigned short)k;
                }
            }
        }
        __syncthreads();
        int2* __restrict__ out = a.pair_out + p0;
        for (int j = threadIdx.x; j < count; j += kBlock)
            out[j] = make_int2(org + (int)s_lk[j], (int)s_sl[j]);
        return;
    }
    // ---- a large bucket: the counts of all its chunks first, then chunk by chunk to the final positions
    for (int k = threadIdx.x; k < kBucketSpan; k += kBlock)
    {
        s_tot[k] = 0u;
        s_fill[k] = 0u;
    }
    __syncthreads();
    for (int j0 = 0; j0 < count; j0 += kBlock)
    { // (warp-aggregated: a cell with thousands of particles would otherwise serialise on one counter)
        const int j = j0 + threadIdx.x;
        const bool valid = j < count;
        const int k = valid ? __ldg(&in[j].x) - org : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (valid && (peers & lanemask_lt()) == 0)
            atomicAdd(&s_tot[k], (uint32_t)__popc(peers));
    }
    __syncthreads();
    bucket_scan_counts(s_tot, s_ws);
    uint32_t* s_ctot = s_sl; // (the staging area is not used on this path)
    for (int c0 = 0; c0 < count; c0 += kTileItems)
    {
        const int cnt = min(kTileItems, count - c0);
        bucket_rank_chunk(in + c0, cnt, org, s_cnt, s_ctot, rank2);
        const unsigned short* my_cnt = s_cnt + warp * kBucketSpan;
#pragma unroll
        for (int r = 0; r < kRoundsPerWarp; r++)
        {
            const int j = wb + r * 32 + lane;
            if (j < cnt)
            {
                const int2 pr = __ldg(in + c0 + j);
                const int k = pr.x - org;
                const uint32_t rk = (r & 1) ? (rank2[r >> 1] >> 16) : (rank2[r >> 1] & 0xffffu);
                a.pair_out[p0 + s_tot[k] + s_fill[k] + my_cnt[k] + rk] = pr;
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < kBucketSpan; k += kBlock)
            s_fill[k] += s_ctot[k];
        __syncthreads();
    }
}

// =========================================================================================================
// per-cell sums over the particles in sorted order (one warp per 256 consecutive sorted positions).
// replaces the reference's  weight scan + prefix differences (dogm.cu:287-288, common.h:25-32, mass_update.cu:76)
// and the five moment scans (dogm.cu:359-377, statistical_moments.cu:58-76); also yields GridCell.start_idx /
// end_idx (particle_to_grid.cu:33-40) and a compact copy of the sorted predicted weights.
//
// Fixed combination order: the piece of a segment that is still open is accumulated lane-locally across the 32-wide
// steps (no shuffles while a step lies inside one cell, the common case once particles have clustered) and
// butterfly-reduced when it closes; segments that begin and end inside one step go through a segmented
// Kogge-Stone scan.  The weight sum is carried in double and rounded once.
// =========================================================================================================
struct Sum6
{
    double s0;
    float s1, s2, s3, s4, s5;
};

__device__ __forceinline__ Sum6 warp_reduce6(Sum6 v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        v.s0 += __shfl_xor_sync(0xffffffffu, v.s0, d);
        v.s1 += __shfl_xor_sync(0xffffffffu, v.s1, d);
        v.s2 += __shfl_xor_sync(0xffffffffu, v.s2, d);
        v.s3 += __shfl_xor_sync(0xffffffffu, v.s3, d);
        v.s4 += __shfl_xor_sync(0xffffffffu, v.s4, d);
        v.s5 += __shfl_xor_sync(0xffffffffu, v.s5, d);
    }
    return v;
}

__device__ __forceinline__ void store_cell_sums(CellSums* sums, int k, const Sum6& v)
{
    float4* p = reinterpret_cast<float4*>(sums + k);
    p[0] = make_float4((float)v.s0, v.s1, v.s2, v.s3);
    p[1] = make_float4(v.s4, v.s5, 0.f, 0.f);
}

__device__ __forceinline__ void store_piece(SegPiece* dst, const Sum6& v)
{
    SegPiece p;
    p.s0 = v.s0;
    p.s1 = v.s1;
    p.s2 = v.s2;
    p.s3 = v.s3;
    p.s4 = v.s4;
    p.s5 = v.s5;
    p.pad = 0;
    *dst = p;
}

__device__ __forceinline__ void sum6_zero(Sum6& v)
{
    v.s0 = 0.0;
    v.s1 = v.s2 = v.s3 = v.s4 = v.s5 = 0.f;
}

__device__ __forceinline__ void sum6_add(Sum6& a, const Sum6& b)
{
    a.s0 += b.s0;
    a.s1 += b.s1;
    a.s2 += b.s2;
    a.s3 += b.s3;
    a.s4 += b.s4;
    a.s5 += b.s5;
}

constexpr int kSegPerLane = kSegChunk / 32; // 8 consecutive sorted positions per lane

// Every lane walks its 8 consecutive positions serially: segments that begin and end inside the lane are finished
// there; the piece in front of the lane's first head ("lead") and the piece behind its last tail ("trail") are
// joined across lanes by ONE segmented warp scan per 256 particles.  Pieces that cross the chunk border go to the
// lead / trail records of the chunk and are joined by k_segfix.  Fixed order: serial inside a lane, Kogge-Stone
// across lanes, chunk order across chunks.
#ifndef DOGM_SEGSUM_MINBLOCKS
#define DOGM_SEGSUM_MINBLOCKS 1
#endif
__device__ __forceinline__ void segsum_chunk(const int2* __restrict__ spair, const PRec* __restrict__ rec, const int n, int* cell_start,
                                             int* cell_end, CellSums* sums, SegPiece* lead, SegPiece* trail, int* flags,
                                             float* __restrict__ sw, const int chunk)
{
    const int lane = threadIdx.x & 31;
    const int base = chunk * kSegChunk;
    if (base >= n)
        return;
    const unsigned full = 0xffffffffu;
    const int p0 = base + lane * kSegPerLane;

    int key[kSegPerLane], slot[kSegPerLane];
    if (p0 + kSegPerLane <= n)
    {
        const int4* q = reinterpret_cast<const int4*>(spair + p0);
#pragma unroll
        for (int j = 0; j < kSegPerLane / 2; j++)
        {
            const int4 t = q[j];
            key[2 * j] = t.x;
            slot[2 * j] = t.y;
            key[2 * j + 1] = t.z;
            slot[2 * j + 1] = t.w;
        }
    }
    else
    {
#pragma unroll
        for (int j = 0; j < kSegPerLane; j++)
        {
            key[j] = -3;
            slot[j] = -1;
            if (p0 + j < n)
            {
                const int2 t = spair[p0 + j];
                key[j] = t.x;
                slot[j] = t.y;
            }
        }
    }
    float4 hi[kSegPerLane]; // (vx, vy, w, -) of the lane's particles: eight independent sector reads in flight
#pragma unroll
    for (int j = 0; j < kSegPerLane; j++)
        // (a plain load: with an evict_last hint on this gather the kernel took 34 instead of 26 us)
        hi[j] = slot[j] >= 0 ? reinterpret_cast<const float4*>(rec + slot[j])[1] : make_float4(0.f, 0.f, 0.f, 0.f);

    int kprev = __shfl_up_sync(full, key[kSegPerLane - 1], 1);
    if (lane == 0)
        kprev = base > 0 ? spair[base - 1].x : -2;
    int knext = __shfl_down_sync(full, key[0], 1);
    if (lane == 31)
        knext = base + kSegChunk < n ? spair[base + kSegChunk].x : -3;

    // ---- serial walk over the lane's positions
    Sum6 cur, leadv;
    sum6_zero(cur);
    sum6_zero(leadv);
    bool seen_head = false, have_lead = false, open = false;
    int lead_key = 0;
    float wout[kSegPerLane];
#pragma unroll
    for (int j = 0; j < kSegPerLane; j++)
    {
        const int p = p0 + j;
        wout[j] = hi[j].z;
        if (p < n)
        {
            const int kj = key[j];
            const int kp = j == 0 ? kprev : key[j - 1];
            const int kn = j == kSegPerLane - 1 ? knext : key[j + 1];
            if (kj != kp)
            {
                cell_start[kj] = p;
                seen_head = true;
            }
            const float w = hi[j].z;
            const float wx = __fmul_rn(w, hi[j].x), wy = __fmul_rn(w, hi[j].y);
            cur.s0 += (double)w;
            cur.s1 += wx;
            cur.s2 += wy;
            cur.s3 += __fmul_rn(wx, hi[j].x);
            cur.s4 += __fmul_rn(wy, hi[j].y);
            cur.s5 += __fmul_rn(wx, hi[j].y);
            open = true;
            if (kj != kn)
            {
                cell_end[kj] = p;
                if (seen_head)
                    store_cell_sums(sums, kj, cur); // began at a head inside this lane
                else
                {
                    leadv = cur; // began in an earlier lane or chunk
                    have_lead = true;
                    lead_key = kj;
                }
                sum6_zero(cur);
                open = false;
            }
        }
    }
    if (p0 + kSegPerLane <= n)
    {
        float4* o = reinterpret_cast<float4*>(sw + p0);
        o[0] = make_float4(wout[0], wout[1], wout[2], wout[3]);
        o[1] = make_float4(wout[4], wout[5], wout[6], wout[7]);
    }
    else
    {
#pragma unroll
        for (int j = 0; j < kSegPerLane; j++)
            if (p0 + j < n)
                sw[p0 + j] = wout[j];
    }

    // ---- join the pieces across lanes: segmented inclusive scan of the trail pieces, segments start at lanes
    // that contain a head (their trail piece begins inside the lane)
    const unsigned fm = __ballot_sync(full, seen_head);
    const unsigned mine = fm & lanemask_le();
    const int start_lane = mine ? (31 - __clz(mine)) : 0;
    Sum6 sc = cur; // zero when the lane's last particle closed its segment
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const double t0 = __shfl_up_sync(full, sc.s0, d);
        const float t1 = __shfl_up_sync(full, sc.s1, d);
        const float t2 = __shfl_up_sync(full, sc.s2, d);
        const float t3 = __shfl_up_sync(full, sc.s3, d);
        const float t4 = __shfl_up_sync(full, sc.s4, d);
        const float t5 = __shfl_up_sync(full, sc.s5, d);
        if (lane - d >= start_lane)
        {
            sc.s0 += t0;
            sc.s1 += t1;
            sc.s2 += t2;
            sc.s3 += t3;
            sc.s4 += t4;
            sc.s5 += t5;
        }
    }
    // what arrives from the lanes in front: the scan value of the previous lane
    Sum6 cin;
    cin.s0 = __shfl_up_sync(full, sc.s0, 1);
    cin.s1 = __shfl_up_sync(full, sc.s1, 1);
    cin.s2 = __shfl_up_sync(full, sc.s2, 1);
    cin.s3 = __shfl_up_sync(full, sc.s3, 1);
    cin.s4 = __shfl_up_sync(full, sc.s4, 1);
    cin.s5 = __shfl_up_sync(full, sc.s5, 1);
    if (lane == 0)
        sum6_zero(cin);
    if (have_lead)
    {
        sum6_add(cin, leadv); // (pieces of earlier lanes) + (this lane's lead piece)
        if (fm & lanemask_lt())
            store_cell_sums(sums, lead_key, cin); // the segment began at a head in an earlier lane of this chunk
        else
            store_piece(&lead[chunk], cin); // it began in an earlier chunk
    }
    // ---- chunk summary
    const int first_is_lead = __shfl_sync(full, (int)(key[0] == kprev), 0); // lane 0 always holds a particle
    const int chunk_open = __shfl_sync(full, (int)open, 31);
    if (lane == 31)
    {
        int f = first_is_lead ? SEG_LEAD : 0;
        if (chunk_open)
        {
            f |= SEG_TRAIL;
            store_piece(&trail[chunk], sc);
            if (fm == 0u)
            {
                f |= SEG_THROUGH;
                store_piece(&lead[chunk], sc);
            }
        }
        flags[chunk] = f;
    }
}

template <bool LOOP>
__global__ void __launch_bounds__(kBlock, DOGM_SEGSUM_MINBLOCKS) k_segsum(const int2* __restrict__ spair, const PRec* __restrict__ rec, int n,
                                                   int* cell_start, int* cell_end, CellSums* sums, SegPiece* lead,
                                                   SegPiece* trail, int* flags, float* __restrict__ sw, const int* n_dev)
{
    pdl_prologue(K_SEGSUM * 2);
    const int chunk0 = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (!LOOP)
    {
        segsum_chunk(spair, rec, n, cell_start, cell_end, sums, lead, trail, flags, sw, chunk0);
        return;
    }
    // device-paced band cycle: the count lives on the device, the grid is the host's estimate, the warps walk over the chunks
    n = *n_dev;
    for (int chunk = chunk0; chunk * kSegChunk < n; chunk += gridDim.x * kWarpsPerBlock)
        segsum_chunk(spair, rec, n, cell_start, cell_end, sums, lead, trail, flags, sw, chunk);
}

// segments that span chunk borders: the chunk holding the head adds up the pieces in chunk order
__global__ void __launch_bounds__(kBlock) k_segfix(const int2* __restrict__ spair, int n_chunks, CellSums* sums,
                                                   const SegPiece* __restrict__ lead, const SegPiece* __restrict__ trail,
                                                   const int* __restrict__ flags, int* zero_word, const int* n_dev)
{
    pdl_prologue(K_SEGFIX * 2);
    if (n_dev)
        n_chunks = (*n_dev + kSegChunk - 1) / kSegChunk;
    if (zero_word && blockIdx.x == 0 && threadIdx.x == 0)
        *zero_word = 0; // the dynamic-cell counter the cell kernel appends to (no memset node between the kernels)
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n_chunks; c += gridDim.x * kBlock) // (one round unless the grid is an estimate)
    {
    const int f = flags[c];
    if (!(f & SEG_TRAIL) || (f & SEG_THROUGH))
        continue;
    SegPiece acc = trail[c];
    int c2 = c + 1;
    while (c2 < n_chunks)
    {
        const SegPiece p = lead[c2];
        acc.s0 += p.s0;
        acc.s1 += p.s1;
        acc.s2 += p.s2;
        acc.s3 += p.s3;
        acc.s4 += p.s4;
        acc.s5 += p.s5;
        if (!(flags[c2] & SEG_THROUGH))
            break;
        c2++;
    }
    const int k = spair[(c + 1) * kSegChunk - 1].x;
    Sum6 v;
    v.s0 = acc.s0;
    v.s1 = acc.s1;
    v.s2 = acc.s2;
    v.s3 = acc.s3;
    v.s4 = acc.s4;
    v.s5 = acc.s5;
    store_cell_sums(sums, k, v);
    }
}

// =========================================================================================================
// persistent-particle weights: updatePersistentParticlesKernel1 + 3 (update_persistent_particles.cu:33-57,78-87)
// with the per-cell constants precomputed by the cell kernel; the over-unit normalisation of
// normalize_weights (mass_update.cu:51-59) is applied on the fly.
// =========================================================================================================
__device__ __forceinline__ float persistent_weight(const float4 c, float w)
{
    if (c.w > 0.0f)
        w = __fdiv_rn(w, c.w);
    const float unnorm = __fmul_rn(c.x, w);
    return __fadd_rn(__fmul_rn(c.y, unnorm), __fmul_rn(c.z, w));
}

__global__ void __launch_bounds__(kBlock) k_weights(const int2* __restrict__ spair, const float* __restrict__ wgt,
                                                    const float4* __restrict__ coef, float* __restrict__ weight_array,
                                                    int n)
{
    pdl_prologue(K_WEIGHTS * 2);
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n)
        return;
    weight_array[i] = persistent_weight(coef[spair[i].x], wgt[i]);
}

// =========================================================================================================
// joint weight CDF (dogm.cu:390-398): reduce-then-scan in double, fixed order.  FUSED: the persistent weights are
// computed here from the sorted predicted weights (k_weights is skipped inside updateGrid); the write pass also
// stores them to weight_array.
// =========================================================================================================
struct CdfArgs
{
    const float* wa;     // weight_array (input when !fused, output when fused)
    float* wa_out;
    const float* bw;     // birth weights
    const int2* spair;
    const float* sw;
    const float4* coef;
    int N, n;
};

template <bool FUSED>
__device__ __forceinline__ float joint_entry(const CdfArgs& a, int i)
{
    if (i < a.N)
        return FUSED ? persistent_weight(a.coef[a.spair[i].x], a.sw[i]) : a.wa[i];
    return i < a.n ? a.bw[i - a.N] : 0.0f;
}

__device__ __forceinline__ float resample_fraction_philox(uint64_t seed, uint32_t slot, uint32_t cycle)
{
    const Philox4 p = philox4x32_10(slot, STAGE_RESAMPLE, cycle, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    return u01_half_open(p.x);
}

// One kernel, one pass: a tile is 4096 consecutive entries (warp w owns entries [512 w, 512 w + 512) and reads them in
// 16 coalesced rounds, entry = 512 w + 32 r + lane, which stay in registers).  Tiles are handed out by an atomic
// ticket, so every tile with a lower number is running or done when a CTA waits for it; a tile publishes its total,
// sums the totals of the tiles before it in a fixed order - the tiles of its own group of
// 256 plus the published prefix of the groups before - and writes its part of the CDF.  The order of every double
// addition is fixed by the tile number alone, so the CDF is identical from run to run.
constexpr int kChainRounds = kCdfTile / kBlock; // 16
constexpr int kChainWarpSpan = kCdfTile / kWarpsPerBlock; // 512
constexpr int kQuadRounds = kChainWarpSpan / 128;          // 4 rounds of 32 lanes x 4 consecutive entries
constexpr int kChainGroup = 256;                // tiles per group
constexpr int kChainFlat = 4;                   // up to kChainFlat * 256 tiles every tile sums all the totals in front of it

struct ChainArgs
{
    CdfArgs c;
    double* cdf;
    double* tile_sum;     // [tiles]       published tile totals (publish_f64 / await_f64)
    double* group_base;   // [groups + 1]  published sum of all tiles of the groups before
    uint32_t* ticket;
    uint32_t ticket_base;
    uint32_t epoch;
    int tiles;
    double* total_out;
    // resampling support: where the first offset of every block of 256 output slots falls in the CDF
    int* res_start;        // [ceil(N / 256)] atomicMin targets, or null (tiles not co-resident / injected offsets)
    double* total_word;    // published total (the last tile's prefix + sum)
    int res_blocks;
    int flat;              // tiles <= kChainFlat * 256: no groups
    int systematic;        // the offsets share one fraction u0 (otherwise the block boundary itself is used, u = 0)
    int noise_injected;
    const float* resample_u;
    uint64_t seed;
    uint32_t cycle;
    const BandCounts* cnt; // device-paced band cycle: particle counts (hence the number of tiles) live on the device
};

// the k-th block of 256 output slots starts at this offset (same expression as resample_offset for slot 256 k)
__device__ __forceinline__ double res_boundary(long long k, double u, double step)
{
    return ((double)(k * 256) + u) * step;
}

// smallest k with res_boundary(k) > x (res_blocks if none)
__device__ __noinline__ long long first_block_behind(double x, double step, double u, double inv, int res_blocks)
{
    if (!(x < __longlong_as_double(0x7ff0000000000000ll)))
        return res_blocks;
    long long k = (long long)ceil(x * inv - u * (1.0 / 256.0));
    k = k < 0 ? 0 : (k > res_blocks ? res_blocks : k);
    while (k < res_blocks && res_boundary(k, u, step) <= x)
        k++;
    while (k > 0 && res_boundary(k - 1, u, step) > x)
        k--;
    return k;
}

// Developer build (-DDOGM_PHASE_TRACE, tools/build_variant.sh): thread 0 of every CTA of k_cdf_chain / k_resample stamps
// %globaltimer at its phase boundaries; dogm_debug_read("phase_chain" / "phase_resample") returns the stamps.
#ifdef DOGM_PHASE_TRACE
constexpr int kPhaseCtas = 4096, kPhaseSlots = 8;
__device__ unsigned long long g_phase[2][kPhaseCtas][kPhaseSlots];
__device__ __forceinline__ void phase_stamp(int which, unsigned cta, int slot, unsigned dep)
{
    if (threadIdx.x == 0 && cta < (unsigned)kPhaseCtas)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer; // %1" : "=l"(t) : "r"(dep) : "memory");
        g_phase[which][cta][slot] = t;
    }
}
#define PHASE_STAMP(which, cta, slot, dep) phase_stamp(which, cta, slot, (unsigned)(dep))
#else
#define PHASE_STAMP(which, cta, slot, dep)
#endif

// as await_f64, but gives up after `polls` attempts (about 50 ns each) and returns -1
__device__ __forceinline__ double await_f64_bounded(const double* p, uint32_t epoch, int polls)
{
    for (int it = 0; it < polls; it++)
    {
        unsigned long long bits;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(bits) : "l"(p) : "memory");
        if ((uint32_t)(bits >> 63) == (epoch & 1u))
            return __longlong_as_double((long long)(bits & 0x7fffffffffffffffull));
        __nanosleep(40);
    }
    return -1.0;
}

template <bool FUSED>
__global__ void __launch_bounds__(kBlock, 4) k_cdf_chain(const ChainArgs a_in)
{
    pdl_prologue(K_CDF_CHAIN * 2);
    ChainArgs a = a_in;
    if (a.cnt)
    {
        a.c.N = a.cnt->n_sort;
        a.c.n = a.cnt->n_sort + a.cnt->B;
        a.tiles = (a.c.n + kCdfTile - 1) / kCdfTile;
        a.flat = a.tiles <= kChainFlat * kBlock ? 1 : 0;
    }
    __shared__ double s_w[kWarpsPerBlock];
    __shared__ double s_red[kWarpsPerBlock];
    __shared__ double s_total, s_step, s_inv, s_u;
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;)
    {
        __syncthreads(); // s_tile / s_w / s_red of the previous tile are no longer read
        if (threadIdx.x == 0)
            s_tile = atomicAdd(a.ticket, 1u) - a.ticket_base;
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= (uint32_t)a.tiles)
            return;
        PHASE_STAMP(0, t, 0, t);
        const int w0 = (int)t * kCdfTile + warp * kChainWarpSpan;
        // Entry layout: in round r (0..3) lane l holds the four consecutive entries w0 + 128 r + 4 l + (0..3), so every
        // load and store is a 16-byte vector and the prefix needs one warp scan per 128 entries (the four entries of a
        // lane are chained serially).  Two rounds are fetched per batch: all (cell, weight) loads of a batch are issued
        // before the first dependent coefficient load, all coefficient loads before the first use.
        float e[kChainRounds];
#pragma unroll
        for (int half = 0; half < 2; half++)
        {
            int cell[8];
            float w[8];
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
            {
                const int i = w0 + (half * 2 + rr) * 128 + lane * 4;
                int* cl = cell + rr * 4;
                float* ww = w + rr * 4;
                if (i + 3 < a.c.N)
                {
                    if (FUSED)
                    {
                        const int4* pp = reinterpret_cast<const int4*>(a.c.spair + i);
                        const int4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
                        const float4 sv = __ldg(reinterpret_cast<const float4*>(a.c.sw + i));
                        cl[0] = p0.x, cl[1] = p0.z, cl[2] = p1.x, cl[3] = p1.z;
                        ww[0] = sv.x, ww[1] = sv.y, ww[2] = sv.z, ww[3] = sv.w;
                    }
                    else
                    {
                        const float4 sv = *reinterpret_cast<const float4*>(a.c.wa + i);
                        cl[0] = cl[1] = cl[2] = cl[3] = -1;
                        ww[0] = sv.x, ww[1] = sv.y, ww[2] = sv.z, ww[3] = sv.w;
                    }
                }
                else
                {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                    {
                        const int k = i + q;
                        cl[q] = -1;
                        if (FUSED)
                        {
                            if (k < a.c.N)
                                cl[q] = __ldg(&a.c.spair[k].x);
                            ww[q] = k < a.c.N ? __ldg(a.c.sw + k) : (k < a.c.n ? __ldg(a.c.bw + (k - a.c.N)) : 0.0f);
                        }
                        else
                        {
                            ww[q] = k < a.c.N ? a.c.wa[k] : (k < a.c.n ? __ldg(a.c.bw + (k - a.c.N)) : 0.0f);
                        }
                    }
                }
            }
            if (FUSED)
            {
                float4 cf[8];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    cf[k] = cell[k] >= 0 ? __ldg(a.c.coef + cell[k]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    e[half * 8 + k] = cell[k] >= 0 ? persistent_weight(cf[k], w[k]) : w[k];
            }
            else
            {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    e[half * 8 + k] = w[k];
            }
        }
        // per round: the prefix in front of this lane's four entries (one warp scan of the lane totals), kept for the
        // write phase; the warp's total is the running carry
        double base[kQuadRounds];
        double carry = 0.0;
#pragma unroll
        for (int r = 0; r < kQuadRounds; r++)
        {
            const double lane_total = (((double)e[4 * r] + (double)e[4 * r + 1]) + (double)e[4 * r + 2]) + (double)e[4 * r + 3];
            double incl = lane_total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const double u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d)
                    incl += u;
            }
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0)
                excl = 0.0;
            base[r] = carry + excl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        PHASE_STAMP(0, t, 1, __float_as_uint(e[15]));
        if (lane == 0)
            s_w[warp] = carry;
        __syncthreads();
        PHASE_STAMP(0, t, 2, 0);
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; w++)
            tot += s_w[w];
        if (threadIdx.x == 0)
            publish_f64(a.tile_sum + t, tot, a.epoch);
        // totals of the tiles before this one, in a fixed order
        const uint32_t group = t / kChainGroup, first = group * kChainGroup;
        double part = 0.0;
        if (a.flat)
        { // up to kChainFlat * 256 tiles: every tile adds up all the words in front of it (kChainFlat per thread, all polls in
          // flight together) - no tile waits for another tile's prefix, only for the totals, which need no look-back
            unsigned long long bits[kChainFlat];
            unsigned need = 0;
#pragma unroll
            for (int k = 0; k < kChainFlat; k++)
                if ((uint32_t)(k * kBlock) + threadIdx.x < t)
                    need |= 1u << k;
            unsigned todo = need;
            while (todo)
            {
#pragma unroll
                for (int k = 0; k < kChainFlat; k++)
                    if (todo & (1u << k))
                        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];"
                                     : "=l"(bits[k])
                                     : "l"(a.tile_sum + k * kBlock + threadIdx.x)
                                     : "memory");
#pragma unroll
                for (int k = 0; k < kChainFlat; k++)
                    if ((todo & (1u << k)) && (uint32_t)(bits[k] >> 63) == (a.epoch & 1u))
                        todo &= ~(1u << k);
                if (todo)
                    __nanosleep(40);
            }
#pragma unroll
            for (int k = 0; k < kChainFlat; k++)
                if (need & (1u << k))
                    part += __longlong_as_double((long long)(bits[k] & 0x7fffffffffffffffull));
        }
        // (a group has at most 256 tiles: one word per thread, all polls in flight together)
        else if (first + threadIdx.x < t)
            part = await_f64(a.tile_sum + first + threadIdx.x, a.epoch);
        else if (group > 0 && first + threadIdx.x == t)
            part = await_f64(a.group_base + group, a.epoch); // prefix of the groups before
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            part += __shfl_xor_sync(0xffffffffu, part, d);
        if (lane == 0)
            s_red[warp] = part;
        __syncthreads();
        PHASE_STAMP(0, t, 3, 0);
        double off = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; w++)
            off += s_red[w];
        if (!a.flat && threadIdx.x == 0 && (t + 1) % kChainGroup == 0 && t + 1 < (uint32_t)a.tiles)
        { // last tile of its group: publish the prefix the next group starts from
            publish_f64(a.group_base + group + 1, off + tot, a.epoch);
        }
        if (threadIdx.x == 0 && t + 1 == (uint32_t)a.tiles)
        { // last tile: the total of the joint weights
            *a.total_out = off + tot;
            if (a.res_start)
                publish_f64(a.total_word, off + tot, a.epoch);
        }
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; w++)
            if (w < warp)
                off += s_w[w];
#pragma unroll
        for (int r = 0; r < kQuadRounds; r++)
        {
            const int i = w0 + r * 128 + lane * 4;
            const double s0 = (double)e[4 * r], s1 = s0 + (double)e[4 * r + 1], s2 = s1 + (double)e[4 * r + 2],
                         s3 = s2 + (double)e[4 * r + 3];
            const double v0 = off + (base[r] + s0), v1 = off + (base[r] + s1), v2 = off + (base[r] + s2), v3 = off + (base[r] + s3);
            if (i + 3 < a.c.n)
            {
                double2* out = reinterpret_cast<double2*>(a.cdf + i);
                __stcs(out, make_double2(v0, v1));
                __stcs(out + 1, make_double2(v2, v3));
            }
            else
            {
                if (i < a.c.n)
                    __stcs(a.cdf + i, v0);
                if (i + 1 < a.c.n)
                    __stcs(a.cdf + i + 1, v1);
                if (i + 2 < a.c.n)
                    __stcs(a.cdf + i + 2, v2);
            }
            if (FUSED)
            {
                if (i + 3 < a.c.N)
                    *reinterpret_cast<float4*>(a.c.wa_out + i) = make_float4(e[4 * r], e[4 * r + 1], e[4 * r + 2], e[4 * r + 3]);
                else
                {
#pragma unroll
                    for (int q = 0; q < 3; q++)
                        if (i + q < a.c.N)
                            a.c.wa_out[i + q] = e[4 * r + q];
                }
            }
        }
        PHASE_STAMP(0, t, 4, 0);
        // third phase (resampling support): which CDF entry is the lower_bound of the first offset of every block of 256
        // output slots.  Needs the total, which the last tile has published long before this tile finished writing.
        if (a.res_start)
        {
            // The wait is bounded: a tile that holds its SM slot for ever while later tiles of this or another stream's
            // launch wait for one could dead-lock the device.  Without the total the tile just makes no claims
            // (k_resample searches for the blocks nobody claimed).
            if (threadIdx.x == 0)
            { // the scalars every thread needs are prepared once (two double divisions and a Philox block)
                const double total = await_f64_bounded(a.total_word, a.epoch, 4096);
                s_total = total;
                if (total > 0.0)
                {
                    const double step = total / (double)a.c.N;
                    s_step = step;
                    s_inv = 1.0 / (256.0 * step);
                    s_u = a.systematic ? (double)(a.noise_injected ? a.resample_u[0] : resample_fraction_philox(a.seed, 0u, a.cycle)) : 0.0;
                }
            }
            __syncthreads();
            PHASE_STAMP(0, t, 5, 0);
            const double total = s_total;
            if (total > 0.0)
            {
                const double step = s_step, inv = s_inv, u = s_u;
                const double kInfD = __longlong_as_double(0x7ff0000000000000ll);
                // the stored value of the entry before the warp's first is only known to a rounding here, so the warp
                // reaches back a little (a claim too many is harmless: atomicMin keeps the true one, which the warp
                // before makes, and k_resample verifies what it reads)
                const double last = w0 == 0 ? -1.0 : off - 1e-9 * fabs(off);
                // k_next: the first block whose boundary lies behind `last`.  A round of 128 entries is looked at closely only
                // when that boundary is not behind its last entry: the lower_bound of the boundary is then the number of
                // entries of the round below it (four ballots), claimed by one lane
                long long k_next = first_block_behind(last, step, u, inv, a.res_blocks);
                double next_off = k_next < a.res_blocks ? res_boundary(k_next, u, step) : kInfD;
#pragma unroll
                for (int r = 0; r < kQuadRounds; r++)
                {
                    const int i = w0 + r * 128 + lane * 4;
                    const double s0 = (double)e[4 * r], s1 = s0 + (double)e[4 * r + 1], s2 = s1 + (double)e[4 * r + 2],
                                 s3 = s2 + (double)e[4 * r + 3];
                    double v[4] = {off + (base[r] + s0), off + (base[r] + s1), off + (base[r] + s2), off + (base[r] + s3)};
                    // (entries behind the end carry weight 0: v[3] of lane 31 is the last valid entry's value)
                    const double hi = __shfl_sync(0xffffffffu, v[3], 31);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (i + q >= a.c.n)
                            v[q] = kInfD;
                    if (w0 + r * 128 < a.c.n)
                    {
                        while (k_next < a.res_blocks && next_off <= hi)
                        {
                            const int below = __popc(__ballot_sync(0xffffffffu, v[0] < next_off)) +
                                              __popc(__ballot_sync(0xffffffffu, v[1] < next_off)) +
                                              __popc(__ballot_sync(0xffffffffu, v[2] < next_off)) +
                                              __popc(__ballot_sync(0xffffffffu, v[3] < next_off));
                            if (lane == 0)
                                atomicMin(a.res_start + k_next, w0 + r * 128 + below);
                            k_next++;
                            next_off = k_next < a.res_blocks ? res_boundary(k_next, u, step) : kInfD;
                        }
                    }
                }
            }
        }
        PHASE_STAMP(0, t, 6, 0);
    }
}

// exclusive scan of up to a few 10^5 block sums by one CTA (thread-serial chunks + one block scan)
__global__ void __launch_bounds__(1024) k_blocksum_scan(const double* __restrict__ in, double* __restrict__ out_excl,
                                                        int n, double* total_out, const int* pub_count, int* pub_host,
                                                        int pub_seq)
{
    pdl_prologue(K_BLOCKSUM_SCAN * 2);
    if (pub_host && threadIdx.x == 0)
    { // the cell kernel in front has completed: its dynamic-cell list (host-mapped) is final; tell the waiting host
        const int found = __ldcg(pub_count);
        *reinterpret_cast<volatile int*>(pub_host) = found;
        __threadfence_system();
        *reinterpret_cast<volatile int*>(pub_host + 1) = pub_seq;
    }
    __shared__ double s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (n + 1023) / 1024;
    const int b0 = min((int)threadIdx.x * per, n), b1 = min(b0 + per, n);
    double local = 0.0;
    for (int b = b0; b < b1; b++)
        local += in[b];
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const double t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += t;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        double wv = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const double t = __shfl_up_sync(0xffffffffu, wv, d);
            if (lane >= d)
                wv += t;
        }
        s_warp[lane] = wv; // inclusive over warps
    }
    __syncthreads();
    const double warp_off = warp > 0 ? s_warp[warp - 1] : 0.0;
    double excl = warp_off + (incl - local);
    for (int b = b0; b < b1; b++)
    {
        const double v = in[b];
        out_excl[b] = excl;
        excl += v;
    }
    if (threadIdx.x == 1023)
        *total_out = s_warp[31];
}

// =========================================================================================================
// resampling: offsets -> ancestor by binary search in the CDF -> gather into the next particle set
// (resampling.cu:34-68; dogm.cu:400-419).  weight_total stays on the device (no joint_max read-back).
// =========================================================================================================
struct ResampleArgs
{
    const double* cdf;
    int n_cdf;
    int N;              // persistent entries of the CDF (the entries behind are birth particles)
    int N_out;          // output slots of this handle (== N without bands)
    int N_glob;         // output slots of the whole grid: step = total / N_glob, new weight = total / N_glob
    long long out_base; // global number of this handle's first output slot (0 without bands)
    double cdf_base;    // joint weight in front of this handle's CDF (0 without bands)
    const PRec* rec;    // persistent particles (records in slot order)
    const int2* spair;  // sorted (cell, slot) pairs: position p of the sorted set is record spair[p].y
    ParticleSet birth;  // birth particles
    ParticleSet dst;   // next population
    int* ancestors;
    const DeviceScalars* scal;
    int* res_start;     // block starts claimed by k_cdf_chain (consumed and reset here), or null
    int mode;           // DOGM_RESAMPLE_*
    int noise_injected; // offsets come from resample_u
    const float* resample_u;
    uint64_t seed;
    uint32_t cycle;
    const BandCounts* cnt; // device-paced band cycle: counts and this band's part of the global draw live on the device
    int* samples;          // bucket sort: cell index of the particle in every stride-th output slot (splitters of the next cycle), or null
    int sample_mask, sample_shift; // stride = 1024 << sample_shift, mask = (1 << sample_shift) - 1
};

// One CTA resamples 1024 consecutive output slots, four per thread.  Their offsets ascend, so all ancestors lie in a short
// window of the CDF behind the ancestor of the CTA's first offset: k_cdf_chain has left that position in res_start (else a
// cooperative search finds it), 2048 CDF entries are staged in shared memory with coalesced loads and every thread
// searches there; the four slot loads and then the four record gathers of a thread are in flight together.  Offsets
// beyond the window (long runs of zero weights) or out of order (caller-supplied fractions that do not ascend) fall back
// to a search in global memory, so the result is lower_bound on the whole CDF in every case.
constexpr int kResWindow = 2048;
constexpr int kResSlides = 6; // further windows a CTA looks at before its open offsets are searched in global memory

__device__ __forceinline__ int lower_bound_f64(const double* __restrict__ cdf, int lo, int hi, double r)
{
    while (lo < hi)
    {
        const int mid = lo + ((hi - lo) >> 1);
        if (cdf[mid] < r)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// the offset of output slot i into the joint weight total (dogm.cu:402-411; BASELINE.json: systematic resampling).
// `u0` (the one fraction of systematic resampling) and `step` = total / N are computed once per CTA.
__device__ __forceinline__ double resample_offset(const ResampleArgs& a, int i, float joint_max, float u0, double step)
{
    if (a.mode == DOGM_RESAMPLE_INJECTED)
        return (double)__fmul_rn(joint_max, a.resample_u[i]);
    const long long gi = a.out_base + i; // slot number in the whole grid
    float u = u0;
    if (a.mode == DOGM_RESAMPLE_STRATIFIED)
        u = a.noise_injected ? a.resample_u[i] : resample_fraction_philox(a.seed, (uint32_t)gi, a.cycle);
    return ((double)gi + (double)u) * step - a.cdf_base;
}

constexpr int kResPer = 4;                     // output slots per thread
constexpr int kResOutputs = kBlock * kResPer;  // output slots per CTA

#ifndef DOGM_RES_MINBLOCKS
#define DOGM_RES_MINBLOCKS 6
#endif
// shared-memory barrier + bulk copy (the copy engine moves the CDF window global -> shared: no registers, one instruction)
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the initialised barrier is visible to the copy engine
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_global, uint32_t bytes, unsigned long long* bar)
{ // 16-byte aligned addresses, size a multiple of 16
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// the 1024 output slots of block `blk`.  BULK: the CDF window is fetched by one bulk copy (cp.async.bulk + mbarrier) issued by
// thread 0 instead of eight register-staged loads per thread (single-shot kernel only: the barrier is used once).
template <bool BULK>
__device__ __forceinline__ void resample_block(const ResampleArgs& a, const int blk)
{
    __shared__ __align__(16) double s_buf[kResWindow + 2]; // up to two entries in front of the window (16-byte aligned source)
    __shared__ __align__(16) int s_anc[kResOutputs];
    __shared__ double s_step, s_first;
    __shared__ float s_u0, s_jm;
    __shared__ __align__(8) unsigned long long s_bar;
    const int out0 = blk * kResOutputs;
    PHASE_STAMP(1, blockIdx.x, 0, 0);
    // The window starts at lower_bound(first offset of the CTA).  k_cdf_chain has normally left that position in
    // res_start (one entry per 256 output slots): every thread reads it; the window entries are fetched straight away,
    // while thread 0 prepares the scalars; the claim is then checked against the CDF (entry before < first offset <= entry).
    int lo0 = 0x7f7f7f7f;
    if (a.res_start)
        lo0 = __ldcg(a.res_start + blk * kResPer);
    const bool claimed = lo0 >= 0 && lo0 < a.n_cdf;
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    // BULK: the copy starts at the even entry lo0a <= lo0 - 1, so that the entry in front of the window comes along
    const int lo0a = (BULK && claimed && lo0 > 0) ? ((lo0 - 1) & ~1) : (claimed ? lo0 : 0);
    const int off = BULK && claimed ? lo0 - lo0a : 0;
    double* s_cdf = s_buf + off; // s_cdf[j] = cdf[lo0 + j]
    double win[BULK ? 1 : kResWindow / kBlock];
    double before = -1.0;
    int got = 0; // BULK: entries the copy delivers
    if (BULK)
    {
        if (claimed)
        {
            got = min(kResWindow + off, a.n_cdf - lo0a);
            if (threadIdx.x == 0)
            {
                const uint32_t bytes = (uint32_t)((got + 1) & ~1) * 8u; // (an odd count reads one entry of padding behind the CDF)
                mbar_init(&s_bar, 1u);
                mbar_expect_tx(&s_bar, bytes);
                bulk_load(s_buf, a.cdf + lo0a, bytes, &s_bar);
            }
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < kResWindow / kBlock; k++)
        {
            const long long j = (long long)lo0 + k * kBlock + (int)threadIdx.x;
            win[BULK ? 0 : k] = (claimed && j < a.n_cdf) ? __ldcg(a.cdf + j) : kInf;
        }
        if (claimed && threadIdx.x == 0 && lo0 > 0)
            before = __ldcg(a.cdf + lo0 - 1);
    }
    if (threadIdx.x == 0)
    {
        const double total = a.scal->weight_total;
        const float jm = (float)total;
        float u0 = 0.0f;
        if (a.mode == DOGM_RESAMPLE_SYSTEMATIC)
            u0 = a.noise_injected ? a.resample_u[0] : resample_fraction_philox(a.seed, 0u, a.cycle);
        const double step = total / (double)a.N_glob;
        s_step = step;
        s_u0 = u0;
        s_jm = jm;
        // a lower limit of the CTA's offsets: its first offset (systematic, injected) or the block boundary (stratified)
        s_first = a.mode == DOGM_RESAMPLE_STRATIFIED ? (double)(a.out_base + out0) * step - a.cdf_base
                                                     : resample_offset(a, out0, jm, u0, step);
    }
    if (BULK)
    {
        if (claimed && threadIdx.x == 0)
            mbar_wait(&s_bar, 0u); // (the other threads are released by the barrier below)
        __syncthreads();
        // the entries behind the end of the CDF read as +inf
        for (int k = got + (int)threadIdx.x; k < kResWindow + 2; k += kBlock)
            s_buf[k] = kInf;
        if (claimed && threadIdx.x == 0)
        {
            before = lo0 > 0 ? s_cdf[-1] : -1.0;
            win[0] = s_cdf[0];
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < kResWindow / kBlock; k++)
            s_cdf[k * kBlock + threadIdx.x] = win[BULK ? 0 : k];
    }
    __syncthreads();
    PHASE_STAMP(1, blockIdx.x, 1, 0);
    const float joint_max = s_jm;
    const double r_first = s_first;
    bool ok = claimed;
    if (claimed && threadIdx.x == 0)
        ok = (lo0 == 0 || before < r_first) && win[0] >= r_first;
    if (a.res_start && threadIdx.x < kResPer)
    { // unclaimed again for the next cycle
        const int claim = blk * kResPer + (int)threadIdx.x;
        if (claim * kBlock < a.N_out)
            a.res_start[claim] = 0x7f7f7f7f;
    }
    const bool have = __syncthreads_and(ok);
    if (!have)
    { // no valid claim: 256-ary search, every thread probes one CDF entry per round
        lo0 = 0;
        int hi0 = a.n_cdf;
        while (lo0 < hi0)
        {
            const int stride = (hi0 - lo0 + kBlock - 1) / kBlock;
            const int q = lo0 + threadIdx.x * stride;
            const int cnt = __syncthreads_count(q < hi0 && a.cdf[q] < r_first); // monotone: the first cnt probes are below
            const int nlo = cnt > 0 ? lo0 + (cnt - 1) * stride + 1 : lo0;
            const long long qn = (long long)lo0 + (long long)cnt * stride;
            hi0 = qn < hi0 ? (int)qn : hi0;
            lo0 = nlo;
        }
        __syncthreads();
        s_cdf = s_buf;
        for (int j = threadIdx.x; j < kResWindow; j += kBlock)
            s_cdf[j] = (lo0 + j < a.n_cdf) ? a.cdf[lo0 + j] : kInf;
        __syncthreads();
    }
    // Ancestors of this thread's four CONSECUTIVE output slots out0 + 4 thread + (0..3): the first by binary search in the
    // window; offsets ascend, so each of the others starts where the one before ended and is normally found within a probe or
    // two.  Where the CDF is dense in entries (the birth particles at its end: many entries of little weight) the CTA's offsets
    // reach beyond the window: the CTA then slides the window on, up to kResSlides times, and the offsets still open are
    // searched there.  Offsets in front of the window (caller-supplied fractions that do not ascend) or still open after the
    // last slide (very long runs of zero weights) are searched in global memory, so that the result is lower_bound on the
    // whole CDF in every case.
    const int ibase = out0 + kResPer * (int)threadIdx.x;
    PHASE_STAMP(1, blockIdx.x, 2, 0);
    double r[kResPer];
    int anc[kResPer];
    unsigned open = 0; // bit j: output j has no ancestor yet
#pragma unroll
    for (int j = 0; j < kResPer; j++)
    {
        const int i = ibase + j;
        r[j] = resample_offset(a, i < a.N_out ? i : a.N_out - 1, joint_max, s_u0, s_step);
        anc[j] = 0;
        if (r[j] < r_first)
            anc[j] = lower_bound_f64(a.cdf, 0, a.n_cdf, r[j]);
        else
            open |= 1u << j;
    }
    for (int slide = 0;; slide++)
    {
        const double w_last = s_cdf[kResWindow - 1];
        int l_prev = 0;
        double r_prev = 0.0;
        bool chained = false; // the output before this one was resolved in this window (l_prev / r_prev are valid)
#pragma unroll
        for (int j = 0; j < kResPer; j++)
        {
            if (!(open & (1u << j)) || !(r[j] <= w_last))
            {
                chained = false;
                continue;
            }
            // every entry in front of the window is below r[j]
            int l = 0, h = kResWindow;
            if (chained && r[j] >= r_prev)
            { // lower_bound(r) >= lower_bound(r_prev): two probes forward, then a search bounded by a short gallop
                l = l_prev;
                h = l;
                if (s_cdf[l] < r[j])
                {
                    l += 1; // l < kResWindow: s_cdf[l_prev] < r <= w_last
                    h = l;
                    if (s_cdf[l] < r[j])
                    {
                        l += 1;
                        h = min(l + 14, kResWindow);
                        if (s_cdf[h - 1] < r[j])
                        {
                            l = h;
                            h = kResWindow;
                        }
                    }
                }
            }
            while (l < h)
            {
                const int mid = l + ((h - l) >> 1);
                if (s_cdf[mid] < r[j])
                    l = mid + 1;
                else
                    h = mid;
            }
            anc[j] = lo0 + l;
            open &= ~(1u << j);
            l_prev = l;
            r_prev = r[j];
            chained = true;
        }
        if (!__syncthreads_or(open != 0)) // (also: nobody reads this window any more)
            break;
        if (slide == kResSlides || lo0 + kResWindow >= a.n_cdf)
        { // (the second condition cannot hold with open offsets: the entries behind the end are +inf; kept as a guard)
#pragma unroll
            for (int j = 0; j < kResPer; j++)
                if (open & (1u << j))
                    anc[j] = lower_bound_f64(a.cdf, min(lo0 + kResWindow, a.n_cdf), a.n_cdf, r[j]);
            break;
        }
        lo0 += kResWindow;
#pragma unroll
        for (int k = 0; k < kResWindow / kBlock; k++)
        {
            const long long e = (long long)lo0 + k * kBlock + (int)threadIdx.x;
            s_cdf[k * kBlock + threadIdx.x] = e < a.n_cdf ? __ldcg(a.cdf + e) : kInf;
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kResPer; j++)
        anc[j] = anc[j] < a.n_cdf ? anc[j] : a.n_cdf - 1;
    // from here on thread t handles the output slots out0 + 256 j + t (every global load and store of a warp is contiguous)
    *reinterpret_cast<int4*>(s_anc + kResPer * threadIdx.x) = make_int4(anc[0], anc[1], anc[2], anc[3]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kResPer; j++)
        anc[j] = s_anc[j * kBlock + threadIdx.x];
    PHASE_STAMP(1, blockIdx.x, 3, anc[kResPer - 1]);
    // gather: all slot loads first, then all record loads
    int slot[kResPer];
#pragma unroll
    for (int j = 0; j < kResPer; j++)
        slot[j] = anc[j] < a.N ? __ldg(&a.spair[anc[j]].y) : -1;
    PHASE_STAMP(1, blockIdx.x, 4, slot[kResPer - 1]);
    float4 st[kResPer];
    int cell[kResPer];
    uint32_t as[kResPer];
#pragma unroll
    for (int j = 0; j < kResPer; j++)
    {
        if (slot[j] >= 0)
        {
            const float4* p = reinterpret_cast<const float4*>(a.rec + slot[j]);
#ifdef DOGM_NO_L2_HINTS
            const float4 rlo = __ldg(p), rhi = __ldg(p + 1);
#else
            const float4 rlo = ld_hint(p, l2_evict_first()), rhi = ld_hint(p + 1, l2_evict_first()); // (their last use)
#endif
            st[j] = make_float4(rlo.x, rlo.y, rhi.x, rhi.y);
            cell[j] = __float_as_int(rlo.z);
            as[j] = __float_as_uint(rlo.w);
        }
        else
        {
            const int b = anc[j] - a.N;
            st[j] = a.birth.state[b];
            cell[j] = a.birth.idx[b];
            as[j] = a.birth.assoc[b];
        }
    }
    PHASE_STAMP(1, blockIdx.x, 5, __float_as_uint(st[kResPer - 1].w) ^ __float_as_uint(st[0].x));
    const float w_new = __fdiv_rn(joint_max, (float)a.N_glob);
    if (a.samples && threadIdx.x == 0 && (blk & a.sample_mask) == 0)
        a.samples[blk >> a.sample_shift] = cell[0]; // output slot out0 = 1024 blk is a multiple of the sample stride
#pragma unroll
    for (int j = 0; j < kResPer; j++)
    {
        const int i = out0 + j * kBlock + (int)threadIdx.x;
        if (i < a.N_out)
        {
            a.ancestors[i] = anc[j];
            a.dst.state[i] = st[j];
            a.dst.idx[i] = cell[j];
            a.dst.assoc[i] = (uint8_t)as[j];
            a.dst.weight[i] = w_new;
        }
    }
    PHASE_STAMP(1, blockIdx.x, 6, 0);
}

template <bool DEV>
__global__ void __launch_bounds__(kBlock, DOGM_RES_MINBLOCKS) k_resample(const ResampleArgs a_in)
{
    pdl_prologue(K_RESAMPLE * 2);
    if (!DEV)
    { // the CTAs at the end of the CDF (birth particles: several window slides) take longest: they go first
        resample_block<true>(a_in, (int)(gridDim.x - 1 - blockIdx.x));
        return;
    }
    // device-paced band cycle: counts and this band's part of the global draw are read from the device; the grid is the host's
    // estimate, the CTAs walk over the blocks of output slots
    ResampleArgs a = a_in;
    a.N = a.cnt->n_sort;
    a.n_cdf = a.cnt->n_sort + a.cnt->B;
    a.N_out = a.cnt->n_out;
    a.out_base = a.cnt->out_base;
    a.cdf_base = a.cnt->cdf_base;
    if (a.n_cdf <= 0)
        return;
    const int blocks = (a.N_out + kResOutputs - 1) / kResOutputs;
    for (int b = blockIdx.x; b < blocks; b += gridDim.x)
    {
        resample_block<false>(a, blocks - 1 - b);
        __syncthreads();
    }
}

// ancestor search on a caller-supplied float CDF: thrust::lower_bound of resampling.cu:45 (+ clamp)
__global__ void __launch_bounds__(kBlock) k_search_f32(const float* __restrict__ cdf, int n_cdf,
                                                       const float* __restrict__ draws, int n_draws, int* out)
{
    pdl_prologue(K_MISC * 2);
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_draws)
        return;
    const float r = draws[i];
    int lo = 0, hi = n_cdf;
    while (lo < hi)
    {
        const int mid = lo + ((hi - lo) >> 1);
        if (cdf[mid] < r)
            lo = mid + 1;
        else
            hi = mid;
    }
    out[i] = lo < n_cdf ? lo : n_cdf - 1;
}

// the noise a Philox-mode cycle consumes, written out for a checker
__global__ void __launch_bounds__(kBlock) k_export_noise(uint64_t seed, uint32_t cycle, int N, int B, float sigma_pos,
                                                         float sigma_vel, float stddev_velocity, float init_max_velocity,
                                                         int resample_mode, float4* predict, float2* birth, float2* init,
                                                         float* resample)
{
    pdl_prologue(K_MISC * 2);
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i < N)
    {
        if (predict)
            predict[i] = predict_noise_philox(seed, (uint32_t)i, cycle, sigma_pos, sigma_vel);
        if (init)
        {
            const Philox4 p = philox4x32_10((uint32_t)i, STAGE_INIT, cycle, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
            const float span = __fsub_rn(init_max_velocity, -init_max_velocity);
            init[i] = make_float2(__fadd_rn(-init_max_velocity, __fmul_rn(span, __fsub_rn(1.0f, u01_open_low(p.x)))),
                                  __fadd_rn(-init_max_velocity, __fmul_rn(span, __fsub_rn(1.0f, u01_open_low(p.y)))));
        }
        if (resample)
            resample[i] = resample_fraction_philox(seed, resample_mode == DOGM_RESAMPLE_STRATIFIED ? (uint32_t)i : 0u, cycle);
    }
    if (i < B && birth)
    {
        const float4 g = philox_normal4(seed, (uint32_t)i, STAGE_BIRTH, cycle);
        birth[i] = make_float2(__fmul_rn(g.x, stddev_velocity), __fmul_rn(g.y, stddev_velocity));
    }
}

// =========================================================================================================
// host-side launchers
// =========================================================================================================

// make the SoA block `pa` current again (read-out / re-entry between stages)
int ensure_soa(dogm_handle* h)
{
    if (h->pa_current || h->N <= 0)
    {
        h->pa_current = true;
        return 0;
    }
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_rec_to_soa, div_up(h->N, kBlock), kBlock, 0, h->rec, h->sorted_valid ? h->spair : nullptr, h->pa.state,
                                                                h->pa.idx, h->pa.weight, h->pa.assoc, h->N);
    h->pa_current = true;
    return (int)cudaGetLastError();
}

// Few tiles (the 250^2 demo size: 74): the cycle is bound by the latency of its dependent launches, and the table scans and the
// pair histogram of the sort are three of them - the scatter kernels then scan the (small) tables themselves and count the next
// pass's digits while they store (k_scatter FUSE).
constexpr int kFuseMaxTiles = 128, kFuseMaxBins = 1024;
bool sort_fused(const dogm_handle* h)
{
    if (h->band.enabled || h->bucket.enabled || h->tiles > kFuseMaxTiles || h->sort_fuse_off)
        return false;
    for (int p = 0; p < h->passes; p++)
        if (h->digit_bins[p] > kFuseMaxBins)
            return false;
    return true;
}

int run_predict(dogm_handle* h, float dt)
{
    if (h->N <= 0)
        return 0;
    int e = ensure_soa(h);
    if (e)
        return e;
    PredictArgs a;
    a.state = h->pa.state;
    a.weight = h->pa.weight;
    a.assoc = h->pa.assoc;
    a.out = h->rec;
    a.key_out = h->key0;
    a.n = h->N;
    a.gs = h->gs;
    a.row0 = h->band.row0;
    a.rows = h->band.rows;
    a.mig = h->band.enabled ? h->band.mig : nullptr;
    a.dt = dt;
    a.p_S = h->params.persistence_prob;
    a.sigma_pos = h->params.stddev_process_noise_position;
    a.sigma_vel = h->params.stddev_process_noise_velocity;
    a.shift_active = (h->shift_particles_pending && h->shift.active) ? 1 : 0;
    a.x_move = h->shift.x_move;
    a.y_move = h->shift.y_move;
    a.noise = h->predict_noise;
    a.seed = h->opts.seed + h->band.salt;
    a.cycle = h->cycle;
    a.hist = h->hist[0];
    a.bins = h->digit_bins[0];
    a.mask = (uint32_t)(h->digit_bins[0] - 1);
    a.leave_cnt[0] = h->band.out_cnt[0];
    a.leave_cnt[1] = h->band.out_cnt[1];
    for (int k = 0; k < 2; k++)
    {
        const bool on = sort_fused(h) && k + 1 < h->passes;
        a.clear[k] = on ? h->hist[k + 1] : nullptr;
        a.clear_n[k] = on ? h->tiles * h->digit_bins[k + 1] : 0;
    }
    const bool bucket = h->bucket.enabled;
    a.bkt_out = h->bucket.bkt;
    a.plan_key = h->bucket.plan_key;
    a.plan_slot = h->bucket.plan_slot;
    a.plan_base = h->bucket.plan_base;
    a.plan_pos = h->bucket.plan_pos;
    a.plan_n = h->bucket.n_spl;
    a.plan_shift = h->bucket.stride_shift;
    if (bucket)
    {
        a.bins = h->bucket.bins;
        a.mask = 0xffffffffu;
        if (!h->bucket.samples_valid)
        {
            LaunchScope ls(h, K_MISC, 0.0);
            launch_chained(h->stream, k_sample_keys, div_up(h->bucket.n_spl, kBlock), kBlock, 0, (const int*)h->pa.idx, h->N,
                           h->bucket.smp_raw, h->bucket.stride_shift);
        }
        BucketPlan pl;
        pl.key = h->bucket.plan_key;
        pl.slot = h->bucket.plan_slot;
        pl.base = h->bucket.plan_base;
        pl.pos = h->bucket.plan_pos;
        pl.org = h->bucket.org;
        pl.n_spl = h->bucket.n_spl;
        pl.stride_shift = h->bucket.stride_shift;
        pl.bins = h->bucket.bins;
        LaunchScope ls(h, K_MISC, 0.0);
        launch_chained(h->stream, k_bucket_plan, 1, kWideBlock, 0, (const int*)h->bucket.smp_raw, pl,
                       a.shift_active ? -(a.x_move + a.gs * a.y_move) : 0, h->C);
    }
    const size_t smem = (size_t)a.bins * sizeof(uint32_t);
    {
        LaunchScope ls(h, K_PREDICT, (bucket ? 59.0 : 57.0) * h->N);
        if (h->opts.noise_mode == DOGM_NOISE_INJECTED)
        {
            if (bucket)
                launch_chained(h->stream, k_predict<true, true>, h->tiles, kWideBlock, smem, a);
            else
                launch_chained(h->stream, k_predict<true, false>, h->tiles, kWideBlock, smem, a);
        }
        else if (bucket)
            launch_chained(h->stream, k_predict<false, true>, h->tiles, kWideBlock, smem, a);
        else
            launch_chained(h->stream, k_predict<false, false>, h->tiles, kWideBlock, smem, a);
    }
    h->shift_particles_pending = false;
    h->hist0_valid = true;
    h->hist0_bucket = bucket;
    h->pa_current = false; // the predicted particles are the records now
    h->rec_valid = true;
    h->sorted_valid = false;
    return (int)cudaGetLastError();
}

int run_assignment(dogm_handle* h)
{
    const int N = h->N; // (device-paced band cycle: an upper bound that sizes the grids; the kernels read the count itself)
    if (N <= 0)
        return 0;
    const int* n_dev = h->band.dev_cnt ? &h->band.dev_cnt->n_sort : nullptr;
    // reinitGridParticleIndices (init.cu:86-93): cell_start is all -1 here: k_init_grid set it and the cell kernel
    // resets every entry it consumes; only when the ranges were never consumed (stage API) a memset is needed
    if (h->ranges_in_soa)
    {
        LaunchScope ls(h, K_MEMSET, 4.0 * h->C);
        cudaMemsetAsync(h->cell_start, 0xff, (size_t)h->C * sizeof(int), h->stream);
    }
    h->ranges_in_soa = true;
    if (!h->hist0_valid)
    {
        const size_t smem = (size_t)h->digit_bins[0] * sizeof(uint32_t);
        const uint32_t mask = (uint32_t)(h->digit_bins[0] - 1);
        LaunchScope ls(h, K_TILE_HIST, 57.0 * N);
        if (h->pa_current || !h->rec_valid)
            launch_chained(h->stream, k_soa_to_rec, h->tiles, kWideBlock, smem, h->pa.state, h->pa.idx, h->pa.weight, h->pa.assoc, h->rec,
                                                                    h->key0, N, h->hist[0], h->digit_bins[0], mask);
        else
            launch_chained(h->stream, k_key_tile_hist, h->tiles, kWideBlock, smem, h->key0, N, h->hist[0], h->digit_bins[0], mask, n_dev);
        h->rec_valid = true;
    }
    const bool fused = sort_fused(h) && !n_dev;
    if (fused && !h->hist0_valid)
    { // (the tables of the later passes are cleared by the prediction kernel; without one in front of this sort, here)
        for (int p = 1; p < h->passes; p++)
            cudaMemsetAsync(h->hist[p], 0, (size_t)h->tiles * h->digit_bins[p] * sizeof(uint32_t), h->stream);
    }
    const bool bucket = h->hist0_valid && h->hist0_bucket;
    if (bucket)
    { // bucket sort: scan of the grouping table, grouping pass (stable, by bucket), one counting sort per bucket
        const int bins = h->bucket.bins;
        {
            LaunchScope ls(h, K_HIST_SCAN, 8.0 * h->tiles * bins);
            launch_chained(h->stream, k_hist_scan, bins / 32, kWideBlock, 0, h->hist[0], h->tiles, bins,
                           (unsigned long long*)h->bin_base[0], ++h->scan_epoch[0], n_dev);
        }
        ScatterArgs a;
        a.key_in = h->key0;
        a.pair_in = nullptr;
        a.pair_out = h->pairs[0];
        a.n = N;
        a.shift = 0;
        a.mask = 0xffffu;
        a.bins = bins;
        a.table = h->hist[0];
        a.n_dev = nullptr;
        a.tiles = h->tiles;
        a.next_hist = nullptr;
        a.next_bins = 1;
        a.next_shift = 0;
        a.digit_in = h->bucket.bkt;
        {
            LaunchScope ls(h, K_SCATTER, 14.0 * N);
            launch_chained(h->stream, k_scatter<2, false>, h->tiles, kBlock, (size_t)bins * kWarpsPerBlock * sizeof(unsigned short), a);
        }
        BucketSortArgs bs;
        bs.pair_in = h->pairs[0];
        bs.pair_out = h->pairs[1];
        bs.start = h->hist[0]; // row 0 of the scanned table: keys in front of every bucket
        bs.org = h->bucket.org;
        bs.bins = bins;
        bs.n = N;
        {
            LaunchScope ls(h, K_SCATTER, 16.0 * N);
            launch_chained(h->stream, k_bucket_sort, bins, kBlock, (size_t)kBucketSmemBytes, bs);
        }
    }
    for (int p = 0; p < (bucket ? 0 : h->passes); p++)
    {
        const int bins = h->digit_bins[p];
        if (fused)
        { // one launch per pass: table scan, ranking, scatter, next pass's tile histogram
            ScatterArgs a;
            a.key_in = h->key0;
            a.pair_in = p > 0 ? h->pairs[(p - 1) & 1] : nullptr;
            a.pair_out = h->pairs[p & 1];
            a.n = N;
            a.shift = h->digit_shift[p];
            a.mask = (uint32_t)(bins - 1);
            a.bins = bins;
            a.table = h->hist[p];
            a.n_dev = nullptr;
            a.digit_in = nullptr;
            a.tiles = h->tiles;
            a.next_hist = p + 1 < h->passes ? h->hist[p + 1] : nullptr;
            a.next_bins = p + 1 < h->passes ? h->digit_bins[p + 1] : 1;
            a.next_shift = p + 1 < h->passes ? h->digit_shift[p + 1] : 0;
            const size_t smem = (size_t)bins * kWarpsPerBlock * sizeof(unsigned short) + 2 * (size_t)bins * sizeof(uint32_t);
            LaunchScope ls(h, K_SCATTER, (p == 0 ? 12.0 : 16.0) * N + 4.0 * h->tiles * bins * h->tiles);
            if (p == 0)
                launch_chained(h->stream, k_scatter<0, false, true>, h->tiles, kBlock, smem, a);
            else
                launch_chained(h->stream, k_scatter<1, false, true>, h->tiles, kBlock, smem, a);
            continue;
        }
        if (p > 0)
        {
            LaunchScope ls(h, K_TILE_HIST, 8.0 * N);
            launch_chained(h->stream, k_pair_tile_hist, h->tiles, kWideBlock, (size_t)bins * sizeof(uint32_t), h->pairs[(p - 1) & 1],
                           N, h->hist[p], bins, h->digit_shift[p], (uint32_t)(bins - 1), n_dev);
        }
        {
            LaunchScope ls(h, K_HIST_SCAN, 8.0 * h->tiles * bins);
            launch_chained(h->stream, k_hist_scan, bins / 32, kWideBlock, 0, h->hist[p], h->tiles, bins,
                           (unsigned long long*)h->bin_base[p], ++h->scan_epoch[p], n_dev);
        }
        ScatterArgs a;
        a.key_in = h->key0;
        a.pair_in = p > 0 ? h->pairs[(p - 1) & 1] : nullptr;
        a.pair_out = h->pairs[p & 1];
        a.n = N;
        a.shift = h->digit_shift[p];
        a.mask = (uint32_t)(bins - 1);
        a.bins = bins;
        a.table = h->hist[p];
        a.n_dev = n_dev;
        a.digit_in = nullptr;
        a.tiles = h->tiles;
        a.next_hist = nullptr;
        a.next_bins = 1;
        a.next_shift = 0;
        const size_t smem = (size_t)bins * kWarpsPerBlock * sizeof(unsigned short);
        {
            LaunchScope ls(h, K_SCATTER, (p == 0 ? 12.0 : 16.0) * N);
            if (n_dev)
            {
                if (p == 0)
                    launch_chained(h->stream, k_scatter<0, true>, h->tiles, kBlock, smem, a);
                else
                    launch_chained(h->stream, k_scatter<1, true>, h->tiles, kBlock, smem, a);
            }
            else if (p == 0)
                launch_chained(h->stream, k_scatter<0, false>, h->tiles, kBlock, smem, a);
            else
                launch_chained(h->stream, k_scatter<1, false>, h->tiles, kBlock, smem, a);
        }
    }
    h->spair = bucket ? h->pairs[1] : h->pairs[(h->passes - 1) & 1];
    h->hist0_valid = false;
    h->hist0_bucket = false;
    h->pa_current = false;
    // per-cell sums + start/end over the sorted order
    {
        LaunchScope ls(h, K_SEGSUM, 44.0 * N);
        if (n_dev)
            launch_chained(h->stream, k_segsum<true>, div_up(h->n_chunks, kWarpsPerBlock), kBlock, 0, h->spair, h->rec, N, h->cell_start,
                           h->cell_end, h->cell_sums, h->seg_lead, h->seg_trail, h->seg_flags, h->sw, n_dev);
        else
            launch_chained(h->stream, k_segsum<false>, div_up(h->n_chunks, kWarpsPerBlock), kBlock, 0, h->spair, h->rec, N, h->cell_start,
                           h->cell_end, h->cell_sums, h->seg_lead, h->seg_trail, h->seg_flags, h->sw, n_dev);
    }
    {
        LaunchScope ls(h, K_SEGFIX, 0.0);
        launch_chained(h->stream, k_segfix, div_up(h->n_chunks, kBlock), kBlock, 0, h->spair, h->n_chunks, h->cell_sums, h->seg_lead,
                                                                      h->seg_trail, h->seg_flags, h->dyn_filter_on ? h->dyn_count : nullptr, n_dev);
    }
    h->sorted_valid = true;
    return (int)cudaGetLastError();
}

int run_persistent_weights(dogm_handle* h, bool defer)
{
    if (h->N <= 0)
        return 0;
    if (!h->sorted_valid)
        return DOGM_ERR_NOT_INITIALIZED;
    h->weights_deferred = defer; // inside updateGrid the CDF kernels compute the weights on the fly
    if (defer)
        return 0;
    LaunchScope ls(h, K_WEIGHTS, 12.0 * h->N);
    launch_chained(h->stream, k_weights, div_up(h->N, kBlock), kBlock, 0, h->spair, h->sw, h->cell_coef, h->weight_array, h->N);
    return (int)cudaGetLastError();
}

int run_blocksum_scan(dogm_handle* h, const double* in, double* out_excl, int n, double* total_out, bool publish_dyn)
{
    const int* pub_count = nullptr;
    int* pub_host = nullptr;
    int pub_seq = 0;
    if (publish_dyn && h->dyn_pub_armed && h->dyn_pub_dev)
    {
        pub_count = h->dyn_count;
        pub_host = h->dyn_pub_dev;
        pub_seq = ++h->dyn_pub_seq;
        h->dyn_pub_armed = false;
        h->dyn_pub_pending = true;
    }
    LaunchScope ls(h, K_BLOCKSUM_SCAN, 16.0 * n);
    launch_chained(h->stream, k_blocksum_scan, 1, 1024, 0, in, out_excl, n, total_out, pub_count, pub_host, pub_seq);
    return (int)cudaGetLastError();
}

int debug_phase_read(int which, void* out_host, size_t bytes)
{
#ifdef DOGM_PHASE_TRACE
    if (which < 0 || which > 1 || bytes > sizeof(g_phase[0]))
        return DOGM_ERR_INVALID_ARGUMENT;
    return (int)cudaMemcpyFromSymbol(out_host, g_phase, bytes, (size_t)which * sizeof(g_phase[0]), cudaMemcpyDeviceToHost);
#else
    (void)which, (void)out_host, (void)bytes;
    return DOGM_ERR_UNSUPPORTED;
#endif
}

int chain_blocks_per_sm()
{
    int a = 0, b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_cdf_chain<true>, kBlock, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_cdf_chain<false>, kBlock, 0);
    const int m = a < b ? a : b;
    return m > 0 ? m : 1;
}

int configure_kernels()
{
    const int max_bins = 1 << kMaxDigitBits;
    const int smem = max_bins * kWarpsPerBlock * (int)sizeof(unsigned short);
    int e = (int)cudaFuncSetAttribute(k_scatter<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = e ? e : (int)cudaFuncSetAttribute(k_scatter<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = e ? e : (int)cudaFuncSetAttribute(k_scatter<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = e ? e : (int)cudaFuncSetAttribute(k_scatter<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = e ? e : (int)cudaFuncSetAttribute(k_scatter<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = e ? e : (int)cudaFuncSetAttribute(k_bucket_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, kBucketSmemBytes);
    e = e ? e : (int)cudaFuncSetAttribute(k_scatter<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = e ? e : (int)cudaFuncSetAttribute(k_scatter<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return e;
}

int run_cdf(dogm_handle* h)
{
    const int N = h->N, n = h->N + h->B;
    if (n <= 0)
        return 0;
    if (N > 0 && !h->sorted_valid && !h->band.dev_cnt)
        return DOGM_ERR_NOT_INITIALIZED; // the CDF reads the sorted weights of dogm_particle_assignment
    CdfArgs ca;
    ca.wa = h->weight_array;
    ca.wa_out = h->weight_array;
    ca.bw = h->birth.weight;
    ca.spair = h->spair;
    ca.sw = h->sw;
    ca.coef = h->cell_coef;
    ca.N = N;
    ca.n = n;
    const bool fused = h->weights_deferred;
    ChainArgs ch;
    ch.c = ca;
    ch.cdf = h->cdf;
    ch.tile_sum = h->tile_sum;
    ch.group_base = h->tile_off;
    ch.ticket = h->chain_flags;
    ch.ticket_base = h->chain_ticket_base;
    ch.epoch = ++h->chain_epoch;
    ch.tiles = h->n_cdf_tiles;
    ch.flat = h->n_cdf_tiles <= kChainFlat * kBlock ? 1 : 0;
    ch.total_out = &h->scal->weight_total;
    const bool resident = h->n_cdf_tiles <= h->chain_capacity;
    ch.res_start = (resident && h->opts.resample_mode != DOGM_RESAMPLE_INJECTED && !h->band.enabled) ? h->res_start : nullptr;
#ifdef DOGM_AB_NO_CLAIMS
    ch.res_start = nullptr;
#endif
    ch.cnt = h->band.dev_cnt;
    if (h->band.enabled)
    { // the number of tiles changes from cycle to cycle, so the parity of the published words cannot alternate: the
      // words are cleared and every launch publishes with parity 1
        // (device-paced cycle: n_cdf_tiles is an estimate, the words are cleared for the band's capacity)
        const size_t clear_tiles = ch.cnt ? (size_t)div_up((long long)h->band.n_cap + h->band.b_cap > 0 ? (long long)h->band.n_cap + h->band.b_cap : 1, kCdfTile)
                                          : (size_t)h->n_cdf_tiles;
        ch.epoch = 1u;
        if (ch.cnt)
        { // (device-paced cycle: cleared by k_band_collect_born, which also resets the ticket - the number of tickets a launch
          //  draws is not known to the host either, so the ticket starts from 0 every cycle)
            h->chain_ticket_base = 0;
            ch.ticket_base = 0;
        }
        else
        {
            cudaMemsetAsync(h->tile_sum, 0, clear_tiles * sizeof(double), h->stream);
            cudaMemsetAsync(h->tile_off, 0, (clear_tiles + 3) * sizeof(double), h->stream);
        }
    }
    ch.total_word = h->tile_off + (h->n_cdf_tiles / kChainGroup + 1);
    ch.res_blocks = div_up(N > 0 ? N : 1, kBlock);
    ch.systematic = h->opts.resample_mode == DOGM_RESAMPLE_SYSTEMATIC ? 1 : 0;
    ch.noise_injected = (h->opts.noise_mode == DOGM_NOISE_INJECTED) ? 1 : 0;
    ch.resample_u = h->resample_u;
    ch.seed = h->opts.seed;
    ch.cycle = h->cycle;
    const int grid = h->n_cdf_tiles < h->chain_capacity ? h->n_cdf_tiles : h->chain_capacity;
    if (!ch.cnt)
        h->chain_ticket_base += (uint32_t)(h->n_cdf_tiles + grid); // every CTA draws one ticket past the end
    {
        LaunchScope ls(h, K_CDF_CHAIN, fused ? 24.0 * N + 12.0 * h->B : 12.0 * n);
        if (fused)
            launch_chained(h->stream, k_cdf_chain<true>, grid, kBlock, 0, ch);
        else
            launch_chained(h->stream, k_cdf_chain<false>, grid, kBlock, 0, ch);
    }
    h->weights_deferred = false;
    return (int)cudaGetLastError();
}

int run_resample_gather(dogm_handle* h)
{
    const int N = h->N, n = h->N + h->B;
    const int n_out = h->band.enabled ? h->band.n_out : N;
    ResampleArgs a;
    a.cdf = h->cdf;
    a.n_cdf = n;
    a.N = N;
    a.N_out = n_out;
    a.N_glob = h->band.enabled ? h->band.n_glob : N;
    a.out_base = h->band.enabled ? h->band.out_base : 0ll;
    a.cdf_base = h->band.enabled ? h->band.cdf_base : 0.0;
    a.rec = h->rec;
    a.spair = h->spair;
    a.birth = h->birth;
    a.dst = h->pa;
    a.ancestors = h->ancestors;
    a.scal = h->scal;
    a.res_start = (h->opts.resample_mode != DOGM_RESAMPLE_INJECTED && !h->band.enabled) ? h->res_start : nullptr;
    a.mode = h->opts.resample_mode;
    a.noise_injected = (h->opts.noise_mode == DOGM_NOISE_INJECTED) ? 1 : 0;
    a.resample_u = h->resample_u;
    a.seed = h->opts.seed;
    a.cycle = h->cycle;
    a.cnt = h->band.dev_cnt;
    a.samples = h->bucket.enabled ? h->bucket.smp_raw : nullptr;
    a.sample_shift = h->bucket.stride_shift - 10;
    a.sample_mask = (1 << a.sample_shift) - 1;
    h->bucket.samples_valid = h->bucket.enabled && n_out > 0 && n > 0;
    if (a.cnt)
    { // device-paced band cycle: the grid is sized for an estimate, the kernel reads what it really has to do and loops
        const int est = h->band.est_out > 0 ? h->band.est_out : 1;
        LaunchScope ls(h, K_RESAMPLE, 69.0 * est);
        launch_chained(h->stream, k_resample<true>, div_up(est, kResOutputs), kBlock, 0, a);
    }
    else if (n_out > 0 && n > 0)
    {
        LaunchScope ls(h, K_RESAMPLE, 69.0 * n_out);
        launch_chained(h->stream, k_resample<false>, div_up(n_out, kResOutputs), kBlock, 0, a);
    }
    // publish (dogm.cu:128): the next population was written straight into particle_array
    h->pa_current = true;
    h->rec_valid = false;
    h->sorted_valid = false;
    h->hist0_valid = false;
    return (int)cudaGetLastError();
}

int run_resampling(dogm_handle* h)
{
    if (h->N <= 0)
        return 0;
    if (!h->sorted_valid)
        return DOGM_ERR_NOT_INITIALIZED; // resampling gathers from the sorted records of dogm_particle_assignment
    int e = run_cdf(h);
    return e ? e : run_resample_gather(h);
}

// band mode: the particles that arrived from the neighbours (records with global coordinates in the two inboxes) become
// the tail of this cycle's record list
// Band mode, outbox: the slots k_predict flagged are compacted into the two send boxes in slot order.
//   k_predict     : per tile of 4096 slots the number of particles leaving through either edge (as doubles: the tile
//                   offsets come from the same single-CTA scan the born masses use)
//   k_outbox_write: rank inside the tile (block scan in slot order) + tile offset -> position in the box
constexpr int kOutboxPer = kTileItems / kBlock; // 16 consecutive slots per thread
__device__ __forceinline__ unsigned long long outbox_thread_counts(const uint8_t* __restrict__ mig, int n, int i0, uint8_t* f)
{
    unsigned long long c = 0;
    if (i0 + kOutboxPer <= n)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(mig + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < kOutboxPer; k++)
            f[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    }
    else
    {
#pragma unroll
        for (int k = 0; k < kOutboxPer; k++)
            f[k] = i0 + k < n ? mig[i0 + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < kOutboxPer; k++)
        c += f[k] == 1 ? 1ull : (f[k] == 2 ? (1ull << 32) : 0ull);
    return c; // low word: lower edge, high word: upper edge
}

// exclusive scans of the two per-tile count arrays (lower / upper edge) in one launch: CTA 0 and CTA 1, same code as k_blocksum_scan
__global__ void __launch_bounds__(1024) k_outbox_scan(const double* __restrict__ in0, const double* __restrict__ in1, double* out0,
                                                      double* out1, int n, double* totals)
{
    pdl_prologue(K_BLOCKSUM_SCAN * 2 + 1);
    const double* __restrict__ in = blockIdx.x ? in1 : in0;
    double* __restrict__ out_excl = blockIdx.x ? out1 : out0;
    __shared__ double s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (n + 1023) / 1024;
    const int b0 = min((int)threadIdx.x * per, n), b1 = min(b0 + per, n);
    double local = 0.0;
    for (int b = b0; b < b1; b++)
        local += in[b];
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const double t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += t;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        double wv = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const double t = __shfl_up_sync(0xffffffffu, wv, d);
            if (lane >= d)
                wv += t;
        }
        s_warp[lane] = wv;
    }
    __syncthreads();
    const double warp_off = warp > 0 ? s_warp[warp - 1] : 0.0;
    double excl = warp_off + (incl - local);
    for (int b = b0; b < b1; b++)
    {
        const double v = in[b];
        out_excl[b] = excl;
        excl += v;
    }
    if (threadIdx.x == 1023)
        totals[blockIdx.x] = s_warp[31];
}

__global__ void __launch_bounds__(kBlock) k_outbox_write(const uint8_t* __restrict__ mig, const PRec* __restrict__ rec, int n,
                                                         const double* __restrict__ off_lo, const double* __restrict__ off_hi,
                                                         PRec* send_lo, PRec* send_hi, int send_cap)
{
    pdl_prologue(K_MISC * 2 + 1);
    __shared__ unsigned long long s_w[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * kTileItems + threadIdx.x * kOutboxPer;
    uint8_t f[kOutboxPer];
    const unsigned long long mine = outbox_thread_counts(mig, n, i0, f);
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += t;
    }
    if (lane == 31)
        s_w[warp] = incl;
    __syncthreads();
    if (mine == 0)
        return;
    unsigned long long before = incl - mine;
    for (int w = 0; w < warp; w++)
        before += s_w[w];
    long long at_lo = (long long)off_lo[blockIdx.x] + (long long)(uint32_t)before;
    long long at_hi = (long long)off_hi[blockIdx.x] + (long long)(uint32_t)(before >> 32);
#pragma unroll
    for (int k = 0; k < kOutboxPer; k++)
    {
        if (f[k] == 0)
            continue;
        const long long at = f[k] == 1 ? at_lo++ : at_hi++;
        if (at < send_cap)
        {
            const float4* src = reinterpret_cast<const float4*>(rec + i0 + k);
            const float4 lo = src[0], hi = src[1];
            float4* so = reinterpret_cast<float4*>((f[k] == 1 ? send_lo : send_hi) + at);
            so[0] = rec_lo(lo.x, lo.y, 0, __float_as_uint(lo.w));
            so[1] = rec_hi(hi.x, hi.y, hi.w); // the weight the ghost gave up
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_band_append(const PRec* __restrict__ in_lo, int n_lo, const PRec* __restrict__ in_hi,
                                                        int n_hi, PRec* __restrict__ rec, int* __restrict__ key0, int n_cur,
                                                        int gs, int row0, int rows)
{
    pdl_prologue(K_MISC * 2);
    const int k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= n_lo + n_hi)
        return;
    const float4* src = reinterpret_cast<const float4*>(k < n_lo ? in_lo + k : in_hi + (k - n_lo));
    const float4 lo = src[0], hi = src[1];
    const int px = min(max(__float2int_rz(lo.x), 0), gs - 1);
    const int py_raw = min(max(__float2int_rz(lo.y), 0), gs - 1) - row0;
    const int py = min(max(py_raw, 0), rows - 1);
    // the sender only knows that the particle left ITS band: one that jumped over this whole band (thin bands, a large
    // ego-motion shift) is kept as a weightless ghost in the edge row instead of adding its weight to a cell it is not in
    float4 hi2 = hi;
    if (py_raw != py)
        hi2.z = 0.0f;
    const int cell = px + gs * py;
    float4* o = reinterpret_cast<float4*>(rec + n_cur + k);
    o[0] = rec_lo(lo.x, lo.y, cell, __float_as_uint(lo.w));
    o[1] = hi2;
    key0[n_cur + k] = cell;
}


// =========================================================================================================
// device-paced band cycle: the bands' messages travel GPU to GPU.  Every exchange is a pair of one-warp kernels: the
// publisher stores this band's share into the other bands' mailboxes (peer stores, value first, then the cycle's
// sequence number behind a system-wide fence); the collector waits until the shares it needs carry the sequence number,
// derives the band's counts from them - in band order, with the very expressions the host-paced phases use - and only
// then lets its dependents launch.  Only the collectors ever spin, one warp each, so several bands on ONE GPU cannot
// starve each other of SM slots; there the orchestrator also enqueues all publishers of an exchange before any of its
// collectors (streams of one GPU may share a hardware queue: a collector must never sit in front of a publisher it waits
// for).  The waits are bounded (a lost peer ends in an error flag, not in a hung GPU).
// =========================================================================================================
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_sys_u32(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys_u32(unsigned int* p, unsigned int v)
{
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kBandWaitNs = 4000000000ull; // 4 s: far beyond any cycle, short enough not to look like a hang

struct BandLink
{
    BandMail* mail[kMaxBands]; // all bands' mailboxes in band order
    int R, me;
    uint32_t seq;
    BandCounts* cnt;
};

// after prediction + outbox compaction: tell the neighbours how many records wait in the send boxes ...
__global__ void __launch_bounds__(32) k_band_publish_sent(BandLink l, const double* __restrict__ out_total, int n_pred, int send_cap)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); // (the collector behind waits for this kernel to finish)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0)
    {
        int err = 0;
        int s_lo = n_pred > 0 ? (int)out_total[0] : 0, s_hi = n_pred > 0 ? (int)out_total[1] : 0;
        if (s_lo > send_cap || s_hi > send_cap)
        {
            err |= BAND_ERR_SEND_OVERFLOW;
            s_lo = min(s_lo, send_cap);
            s_hi = min(s_hi, send_cap);
        }
        l.cnt->sent_lo = s_lo;
        l.cnt->sent_hi = s_hi;
        l.cnt->err = err; // (first kernel of the cycle that writes the flags)
        __threadfence_system(); // the send boxes (written by the kernel in front) are visible before the counts
        if (l.me > 0)
            st_sys_u64(&l.mail[l.me - 1]->sent_from_hi, ((unsigned long long)l.seq << 32) | (unsigned)s_lo);
        if (l.me + 1 < l.R)
            st_sys_u64(&l.mail[l.me + 1]->sent_from_lo, ((unsigned long long)l.seq << 32) | (unsigned)s_hi);
    }
}

// ... and learn how many they sent
__global__ void __launch_bounds__(32) k_band_collect_sent(BandLink l, int n_pred, int n_cap)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0)
    {
        int err = 0;
        int n_lo = 0, n_hi = 0;
        const unsigned long long t0 = global_ns();
        const BandMail* mine = l.mail[l.me];
        bool need_lo = l.me > 0, need_hi = l.me + 1 < l.R;
        while (need_lo || need_hi)
        {
            if (need_lo)
            {
                const unsigned long long w = ld_sys_u64(&mine->sent_from_lo);
                if ((uint32_t)(w >> 32) == l.seq)
                {
                    n_lo = (int)(uint32_t)w;
                    need_lo = false;
                }
            }
            if (need_hi)
            {
                const unsigned long long w = ld_sys_u64(&mine->sent_from_hi);
                if ((uint32_t)(w >> 32) == l.seq)
                {
                    n_hi = (int)(uint32_t)w;
                    need_hi = false;
                }
            }
            if ((need_lo || need_hi) && global_ns() - t0 > kBandWaitNs)
            {
                err |= BAND_ERR_TIMEOUT;
                break;
            }
        }
        __threadfence_system();
        int n_sort = n_pred + n_lo + n_hi;
        if (n_sort > n_cap)
        {
            err |= BAND_ERR_PARTICLE_CAP;
            n_hi = max(0, min(n_hi, n_cap - n_pred - n_lo));
            n_lo = max(0, min(n_lo, n_cap - n_pred));
            n_sort = n_pred + n_lo + n_hi;
        }
        l.cnt->n_lo = n_lo;
        l.cnt->n_hi = n_hi;
        l.cnt->n_sort = n_sort;
        if (err)
            l.cnt->err |= err;
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// the neighbours' records (read from their send boxes over NVLink) join this band's records; the rows of the neighbours'
// previous free masses an ego-motion shift pulls in are copied into the halo buffers
__global__ void __launch_bounds__(kBlock) k_band_pull(const BandCounts* __restrict__ cnt, const PRec* in_lo, const PRec* in_hi,
                                                      PRec* __restrict__ rec, int* __restrict__ key0, int n_pred, int gs, int row0,
                                                      int rows, const float* edge_lo, const float* edge_hi, float* halo_lo,
                                                      float* halo_hi, int halo_elems)
{
    pdl_prologue(K_MISC * 2);
    const int n_lo = cnt->n_lo, n_hi = cnt->n_hi;
    const int stride = gridDim.x * kBlock;
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < n_lo + n_hi; k += stride)
    {
        const float4* src = reinterpret_cast<const float4*>(k < n_lo ? in_lo + k : in_hi + (k - n_lo));
        const float4 lo = __ldcg(src), hi = __ldcg(src + 1);
        const int px = min(max(__float2int_rz(lo.x), 0), gs - 1);
        const int py_raw = min(max(__float2int_rz(lo.y), 0), gs - 1) - row0;
        const int py = min(max(py_raw, 0), rows - 1);
        float4 hi2 = hi;
        if (py_raw != py)
            hi2.z = 0.0f; // (see k_band_append)
        const int cell = px + gs * py;
        float4* o = reinterpret_cast<float4*>(rec + n_pred + k);
        o[0] = rec_lo(lo.x, lo.y, cell, __float_as_uint(lo.w));
        o[1] = hi2;
        key0[n_pred + k] = cell;
    }
    const int quads = halo_elems / 4; // (a row length that is no multiple of 4 leaves a tail for the scalar loop)
    if (edge_lo)
    {
        for (int k = blockIdx.x * kBlock + threadIdx.x; k < quads; k += stride)
            reinterpret_cast<float4*>(halo_lo)[k] = __ldcg(reinterpret_cast<const float4*>(edge_lo) + k);
        for (int k = quads * 4 + blockIdx.x * kBlock + threadIdx.x; k < halo_elems; k += stride)
            halo_lo[k] = __ldcg(edge_lo + k);
    }
    if (edge_hi)
    {
        for (int k = blockIdx.x * kBlock + threadIdx.x; k < quads; k += stride)
            reinterpret_cast<float4*>(halo_hi)[k] = __ldcg(reinterpret_cast<const float4*>(edge_hi) + k);
        for (int k = quads * 4 + blockIdx.x * kBlock + threadIdx.x; k < halo_elems; k += stride)
            halo_hi[k] = __ldcg(edge_hi + k);
    }
}

// publishes this band's entry of a normaliser (which = 0: born mass, 1: joint weight) into every band's mailbox: lane k serves
// band k - value, system-wide fence, then the tag
__global__ void __launch_bounds__(32) k_band_publish_share(BandLink l, int which, const DeviceScalars* scal)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool any = which == 0 || l.cnt->n_sort + l.cnt->B > 0; // (no CDF tile ran: nobody wrote the total)
    const double share = which ? (any ? scal->weight_total : 0.0) : scal->born_total;
    const int lane = threadIdx.x;
    if (lane == 0)
    {
        if (which)
            l.cnt->weight_local = share;
        else
            l.cnt->born_local = share;
    }
    if (lane < l.R)
    {
        BandMail* m = l.mail[lane];
        st_sys_u64(reinterpret_cast<unsigned long long*>(which ? &m->weight[l.me] : &m->born[l.me]),
                   (unsigned long long)__double_as_longlong(share));
    }
    __threadfence_system();
    if (lane < l.R)
    {
        BandMail* m = l.mail[lane];
        st_sys_u32(which ? &m->weight_seq[l.me] : &m->born_seq[l.me], l.seq);
    }
}

// waits for all bands' entries of a normaliser in this band's own mailbox; sums in band order: every band computes the same doubles
__device__ __forceinline__ int band_collect(const BandLink& l, int which, double* before, double* total)
{
    const int lane = threadIdx.x;
    int err = 0;
    double v = 0.0;
    if (lane < l.R)
    {
        const BandMail* mine = l.mail[l.me];
        const unsigned long long t0 = global_ns();
        while (ld_sys_u32(which ? &mine->weight_seq[lane] : &mine->born_seq[lane]) != l.seq)
            if (global_ns() - t0 > kBandWaitNs)
            {
                err = BAND_ERR_TIMEOUT;
                break;
            }
        __threadfence_system();
        v = __longlong_as_double((long long)ld_sys_u64(reinterpret_cast<const unsigned long long*>(which ? &mine->weight[lane] : &mine->born[lane])));
    }
    err = __reduce_or_sync(0xffffffffu, err);
    double acc = 0.0, b = 0.0;
    for (int k = 0; k < l.R; k++)
    {
        const double vk = __shfl_sync(0xffffffffu, v, k);
        if (k == l.me)
            b = acc;
        acc = acc + vk;
    }
    *before = b;
    *total = acc;
    return err;
}

// after the occupancy update: born mass of the whole grid, this band's birth slots
// (the other warps clear the words of the CDF chain meanwhile: the number of its tiles changes from cycle to cycle, see run_cdf)
__global__ void __launch_bounds__(kBlock) k_band_collect_born(BandLink l, DeviceScalars* scal, int b_glob, int b_cap, double* clear_a,
                                                              double* clear_b, int clear_n, uint32_t* ticket)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x >= 32)
    {
        for (int k = threadIdx.x - 32; k < clear_n; k += kBlock - 32)
        {
            clear_a[k] = 0.0;
            clear_b[k] = 0.0;
        }
        if (threadIdx.x == 32)
            *ticket = 0u;
    }
    double before = 0.0, total = 0.0;
    int err = 0;
    if (threadIdx.x < 32)
        err = band_collect(l, 0, &before, &total);
    if (threadIdx.x == 0)
    {
        const double local = l.cnt->born_local;
        int first, past;
        band_slot_range(before, local, total, b_glob, &first, &past);
        int b = past - first;
        if (b > b_cap)
        {
            err |= BAND_ERR_BIRTH_CAP;
            b = b_cap;
        }
        l.cnt->born_base = before;
        l.cnt->born_total = total;
        l.cnt->birth_slot_base = first;
        l.cnt->B = b;
        if (err)
            l.cnt->err |= err;
        scal->born_total = total;
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// after the joint CDF: joint weight of the whole grid, the output slots this band draws
__global__ void __launch_bounds__(32) k_band_collect_weight(BandLink l, DeviceScalars* scal, uint64_t seed, uint32_t cycle, int systematic,
                                                            int n_glob, int n_cap)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    double before, total;
    int err = band_collect(l, 1, &before, &total);
    if (threadIdx.x == 0)
    {
        const double local = l.cnt->weight_local;
        long long i_lo, i_hi;
        band_output_range(seed, cycle, systematic != 0, (long long)n_glob, before, local, total, &i_lo, &i_hi);
        long long n_out = i_hi - i_lo;
        if (n_out > n_cap)
        {
            err |= BAND_ERR_PARTICLE_CAP;
            n_out = n_cap;
        }
        l.cnt->cdf_base = before;
        l.cnt->weight_total = total;
        l.cnt->out_base = i_lo;
        l.cnt->n_out = (int)n_out;
        if (err)
            l.cnt->err |= err;
        scal->weight_total = total;
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

static BandLink make_link(dogm_handle* h)
{
    BandLink l;
    for (int k = 0; k < kMaxBands; k++)
        l.mail[k] = h->band.peer_mail[k];
    l.R = h->band.group_size;
    l.me = h->band.group_rank;
    l.seq = h->band.seq;
    l.cnt = h->band.cnt;
    return l;
}

int run_band_publish_sent(dogm_handle* h)
{
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_band_publish_sent, 1, 32, 0, make_link(h), h->band.out_total, h->N, h->band.send_cap);
    return (int)cudaGetLastError();
}

int run_band_collect_sent(dogm_handle* h, int n_pred)
{
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_band_collect_sent, 1, 32, 0, make_link(h), n_pred, h->band.n_cap);
    return (int)cudaGetLastError();
}

int run_band_pull(dogm_handle* h, int n_pred, const void* outbox_lo, const void* outbox_hi, const float* edge_lo, const float* edge_hi)
{
    const int halo_elems = h->band.halo_rows * h->gs;
    const long long work = (long long)2 * h->band.send_cap > halo_elems / 4 ? (long long)2 * h->band.send_cap : halo_elems / 4;
    int grid = div_up(work > 0 ? work : 1, kBlock);
    if (grid > 4 * h->sm_count)
        grid = 4 * h->sm_count;
    LaunchScope ls(h, K_MISC, 68.0 * h->band.send_cap);
    launch_chained(h->stream, k_band_pull, grid, kBlock, 0, h->band.cnt, (const PRec*)outbox_lo, (const PRec*)outbox_hi, h->rec, h->key0,
                   n_pred, h->gs, h->band.row0, h->band.rows, edge_lo, edge_hi, h->band.halo[0], h->band.halo[1], halo_elems);
    return (int)cudaGetLastError();
}

int run_band_publish_share(dogm_handle* h, int which)
{
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_band_publish_share, 1, 32, 0, make_link(h), which, (const DeviceScalars*)h->scal);
    return (int)cudaGetLastError();
}

int run_band_collect_born(dogm_handle* h)
{
    LaunchScope ls(h, K_MISC, 0.0);
    // the CDF chain's words for the band's capacity (+ 3 as in run_cdf), cleared here instead of by three memset nodes
    const int clear_n = div_up((long long)h->band.n_cap + h->band.b_cap > 0 ? (long long)h->band.n_cap + h->band.b_cap : 1, kCdfTile) + 3;
    launch_chained(h->stream, k_band_collect_born, 1, kBlock, 0, make_link(h), h->scal, h->band.b_glob, h->band.b_cap, h->tile_sum,
                   h->tile_off, clear_n, h->chain_flags);
    return (int)cudaGetLastError();
}

int run_band_collect_weight(dogm_handle* h)
{
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_band_collect_weight, 1, 32, 0, make_link(h), h->scal, h->opts.seed, h->cycle,
                   h->opts.resample_mode == DOGM_RESAMPLE_SYSTEMATIC ? 1 : 0, h->band.n_glob, h->band.n_cap);
    return (int)cudaGetLastError();
}

int run_band_outbox(dogm_handle* h)
{
    const int n = h->N;
    const int tiles = div_up(n > 0 ? n : 1, kTileItems);
    // (the per-tile counts of the leaving particles were written by k_predict)
    {
        LaunchScope ls(h, K_BLOCKSUM_SCAN, 32.0 * tiles);
        launch_chained(h->stream, k_outbox_scan, 2, 1024, 0, (const double*)h->band.out_cnt[0], (const double*)h->band.out_cnt[1],
                       h->band.out_off[0], h->band.out_off[1], tiles, h->band.out_total);
    }
    int e = (int)cudaGetLastError();
    if (e)
        return e;
    {
        LaunchScope ls(h, K_MISC, 1.0 * n);
        launch_chained(h->stream, k_outbox_write, tiles, kBlock, 0, h->band.mig, h->rec, n, h->band.out_off[0], h->band.out_off[1],
                       h->band.send[0], h->band.send[1], h->band.send_cap);
    }
    return (int)cudaGetLastError();
}

int run_band_append(dogm_handle* h, int n_from_lo, int n_from_hi)
{
    const int add = n_from_lo + n_from_hi;
    if (add <= 0)
        return 0;
    if (h->N + add > h->band.n_cap)
        return DOGM_ERR_INVALID_ARGUMENT;
    {
        LaunchScope ls(h, K_MISC, 68.0 * add);
        launch_chained(h->stream, k_band_append, div_up(add, kBlock), kBlock, 0, h->band.recv[0], n_from_lo, h->band.recv[1], n_from_hi,
                       h->rec, h->key0, h->N, h->gs, h->band.row0, h->band.rows);
    }
    set_particle_counts(h, h->N + add, h->B);
    h->hist0_valid = false; // the histogram of the first sort pass has to include the new records
    return (int)cudaGetLastError();
}

int run_search_ancestors_f32(dogm_handle* h, const float* d_cdf, int n_cdf, const float* d_draws, int n_draws, int* d_out)
{
    if (n_draws <= 0)
        return 0;
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_search_f32, div_up(n_draws, kBlock), kBlock, 0, d_cdf, n_cdf, d_draws, n_draws, d_out);
    return (int)cudaGetLastError();
}

int run_export_noise(dogm_handle* h, uint32_t cycle, float4* d_predict, float2* d_birth, float2* d_init,
                     float* d_resample)
{
    const int m = h->N > h->B ? h->N : h->B;
    if (m <= 0)
        return 0;
    LaunchScope ls(h, K_MISC, 0.0);
    launch_chained(h->stream, k_export_noise, div_up(m, kBlock), kBlock, 0, 
        h->opts.seed, cycle, h->N, h->B, h->params.stddev_process_noise_position, h->params.stddev_process_noise_velocity,
        h->params.stddev_velocity, h->params.init_max_velocity, h->opts.resample_mode, d_predict, d_birth, d_init,
        d_resample);
    return (int)cudaGetLastError();
}

int trace_bind_particles(unsigned long long* p)
{
    return (int)cudaMemcpyToSymbol(c_trace, &p, sizeof(p));
}

} // namespace dogm_b200
