// C ABI of libdogm_b200 (include/dogm_b200.h): host orchestration of the DOGM cycle.
// One stream, every buffer allocated once in dogm_create, no allocation / host read-back inside a cycle
// (the reference does ~26 cudaMalloc/cudaFree pairs and 3 blocking scalar copies per cycle).
#include "dogm_internal.cuh"

#include <atomic>
#include "philox.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace dogm_b200;

namespace dogm_b200
{

void launch_begin(dogm_handle* h, int id, double algorithmic_bytes)
{
    if (id != K_MEMSET)
        h->launch_count++;
    h->last_bytes[id] = algorithmic_bytes;
    if (!h->timing)
        return;
    h->acc_bytes[id] += algorithmic_bytes;
    TimedLaunch t;
    t.id = id;
    for (int k = 0; k < 2; k++)
    {
        cudaEvent_t e;
        if (!h->event_pool.empty())
        {
            e = h->event_pool.back();
            h->event_pool.pop_back();
        }
        else
        {
            cudaEventCreate(&e);
        }
        if (k == 0)
            t.e0 = e;
        else
            t.e1 = e;
    }
    cudaEventRecord(t.e0, h->stream);
    h->timed.push_back(t);
}

void launch_end(dogm_handle* h, int id)
{
    (void)id;
    if (!h->timing)
        return;
    cudaEventRecord(h->timed.back().e1, h->stream);
}

} // namespace dogm_b200

// ---------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------
static int alloc_zero(void** p, size_t bytes)
{
    if (bytes == 0)
        bytes = 16;
    DOGM_CHECK(cudaMalloc(p, bytes));
    DOGM_CHECK(cudaMemset(*p, 0, bytes));
    return 0;
}

static void plan_sort(dogm_handle* h)
{
    int bits = 1;
    while ((1ll << bits) < (long long)h->C)
        bits++;
    h->key_bits = bits;
    // (DOGM_B200_MAX_DIGIT_BITS, tests: narrower digits make a small grid run the three-pass plan of the 16384^2 grid)
    int max_bits = kMaxDigitBits;
    if (const char* mb = getenv("DOGM_B200_MAX_DIGIT_BITS"))
        if (atoi(mb) >= 5 && atoi(mb) <= kMaxDigitBits && (bits + atoi(mb) - 1) / atoi(mb) <= kMaxPasses)
            max_bits = atoi(mb);
    int passes = (bits + max_bits - 1) / max_bits;
    if (passes < 1)
        passes = 1;
    h->passes = passes;
    int shift = 0;
    for (int p = 0; p < passes; p++)
    {
        int d = (bits - shift + (passes - p) - 1) / (passes - p); // spread the remaining bits evenly
        if (d < 5)
            d = 5;
        h->digit_shift[p] = shift;
        h->digit_bins[p] = 1 << d;
        shift += d;
    }
    h->tiles = div_up(h->N > 0 ? h->N : 1, kTileItems);
}

static int copy_in(void* dst_device, const void* src, size_t bytes, int on_device, cudaStream_t s)
{
    return (int)cudaMemcpyAsync(dst_device, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s);
}

// ---------------------------------------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------------------------------------
static int create_impl(const dogm_params* params, const dogm_band_config* band, dogm_handle** out)
{
    if (!params || !out)
        return DOGM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!(params->resolution > 0.0f) || params->particle_count < 0 || params->new_born_particle_count < 0)
        return DOGM_ERR_INVALID_ARGUMENT;
    const int gs = (int)(params->size / params->resolution); // dogm.cu:34
    const int rows = band ? band->rows : gs;
    if (gs <= 0 || rows <= 0 || (long long)gs * rows > (1ll << 30))
        return DOGM_ERR_INVALID_ARGUMENT;
    if (band && (band->row0 < 0 || band->row0 + band->rows > gs || band->particle_capacity < 0 || band->birth_capacity < 0 ||
                 band->exchange_capacity < 0 || band->halo_rows < 0))
        return DOGM_ERR_INVALID_ARGUMENT;
    const long long n_cap = band ? band->particle_capacity : params->particle_count;
    const long long b_cap = band ? band->birth_capacity : params->new_born_particle_count;
    if (n_cap + b_cap >= (1ll << 31) || (long long)params->particle_count + params->new_born_particle_count >= (1ll << 31))
        return DOGM_ERR_INVALID_ARGUMENT;

    int device = 0;
    DOGM_CHECK(cudaGetDevice(&device)); // binds to the current device, dogm.cu:39-43
    int sm_count = 0;
    DOGM_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    DOGM_CHECK((cudaError_t)configure_kernels());

    dogm_handle* h = new (std::nothrow) dogm_handle();
    if (!h)
        return (int)cudaErrorMemoryAllocation;
    h->params = *params;
    h->opts.seed = 123456ull; // the reference's seed, dogm.cu:101
    h->opts.resample_mode = DOGM_RESAMPLE_SYSTEMATIC;
    h->opts.noise_mode = DOGM_NOISE_PHILOX;
    h->gs = gs;
    h->C = gs * rows;
    // sized for the capacities; without bands these are the particle counts themselves
    h->N = (int)n_cap;
    h->B = (int)b_cap;
    h->band.enabled = band ? 1 : 0;
    h->band.G = gs;
    h->band.row0 = band ? band->row0 : 0;
    h->band.rows = rows;
    h->band.n_cap = (int)n_cap;
    h->band.b_cap = (int)b_cap;
    h->band.n_glob = params->particle_count;
    h->band.b_glob = params->new_born_particle_count;
    h->band.salt = band ? band->seed_salt : 0ull;
    h->band.born_base = 0.0;
    h->band.birth_slot_base = 0;
    h->band.cdf_base = 0.0;
    h->band.out_base = 0;
    h->band.n_out = 0;
    h->band.send[0] = h->band.send[1] = h->band.recv[0] = h->band.recv[1] = nullptr;
    h->band.mig = nullptr;
    h->band.pin = nullptr;
    h->band.out_cnt[0] = h->band.out_cnt[1] = h->band.out_off[0] = h->band.out_off[1] = h->band.out_total = nullptr;
    h->band.send_cap = band ? band->exchange_capacity : 0;
    h->band.halo[0] = h->band.halo[1] = nullptr;
    h->band.halo_rows = band ? band->halo_rows : 0;
    h->band.cnt = nullptr;
    h->band.cnt_host = nullptr;
    h->band.mail = nullptr;
    for (int k = 0; k < kMaxBands; k++)
        h->band.peer_mail[k] = nullptr;
    h->band.group_size = 0;
    h->band.group_rank = 0;
    h->band.dev_cnt = nullptr;
    h->band.seq = 0;
    h->band.n_pred = 0;
    h->band.profile = false;
    for (int k = 0; k < 8; k++)
        h->band.stage_ev[k] = nullptr;
    for (int k = 0; k < 7; k++)
        h->band.stage_ms[k] = 0.0f;
    h->band.est_recv = 65536;
    h->band.est_birth = band ? band->birth_capacity : 0;
    h->band.est_out = band ? band->particle_capacity : 0;
    {
        // developer / test switch: fixed (deliberately wrong) launch-size estimates, to exercise the kernels' loops
        const char* ef = getenv("DOGM_B200_BAND_EST");
        h->band.est_forced = ef ? atoi(ef) : 0;
        if (h->band.est_forced > 0)
            h->band.est_recv = h->band.est_birth = h->band.est_out = h->band.est_forced;
    }
    h->device = device;
    h->sm_count = sm_count;
    h->first_pose_received = false;
    h->first_measurement_received = false;
    h->position_x = h->position_y = h->yaw = 0.0f;
    h->shift.active = 0;
    h->shift.x_move = h->shift.y_move = 0;
    h->shift_particles_pending = h->shift_grid_pending = false;
    h->cycle = 0;
    h->hist0_valid = false;
    h->grid_alt = nullptr;
    h->copy_stream = nullptr;
    h->grid_copy_busy[0] = h->grid_copy_busy[1] = false;
    h->grid_swap_pending = false;
    h->launch_count = 0;
    h->timing = false;
    h->timer_ready = false;
    h->predict_noise = nullptr;
    h->birth_noise = nullptr;
    h->init_velocity = nullptr;
    h->resample_u = nullptr;
    h->dyn_buf = nullptr;
    h->dyn_capacity = 0;
    h->dyn_count = nullptr;
    h->dyn_count_host = nullptr;
    h->dyn_filter_on = h->dyn_list_valid = false;
    h->dyn_filter_capacity = 0;
    h->dyn_mapped_host = h->dyn_mapped_dev = nullptr;
    h->cell_kernel_done = true;
    h->lazy_meas.pending = false;
    h->lazy_meas.geom = nullptr;
    h->lazy_meas.polar = nullptr;
    h->lazy_meas.K = h->lazy_meas.H = 0;
    h->dyn_pub_host = h->dyn_pub_dev = nullptr;
    h->dyn_pub_seq = 0;
    h->dyn_pub_armed = h->dyn_pub_pending = false;
    for (int k = 0; k < K_COUNT; k++)
    {
        h->acc_ms[k] = 0.0;
        h->acc_launches[k] = 0;
        h->last_bytes[k] = 0.0;
        h->acc_bytes[k] = 0.0;
    }
    plan_sort(h);
    h->n_cell_blocks = div_up(h->C, kCellBlock);
    h->n_blk_groups = div_up(h->n_cell_blocks, 256);
    h->birth_epoch = 0;
    h->n_chunks = div_up(h->N > 0 ? h->N : 1, kSegChunk);
    h->n_cdf_tiles = div_up((long long)h->N + h->B > 0 ? (long long)h->N + h->B : 1, kCdfTile);

    // from here on every failure goes through dogm_destroy(h) (cudaFree / cudaFreeHost accept the null pointers of a
    // half-built handle); the first error code is the one returned
    int e = 0;
    auto keep_first = [&e](int code) {
        if (e == 0)
            e = code;
    };
    h->stream = nullptr;
    keep_first(check_cuda(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), __FILE__, __LINE__));
    if (e)
    {
        delete h;
        return e;
    }

    const size_t C = (size_t)h->C, N = (size_t)h->N, B = (size_t)h->B;
    void* blk = nullptr;
    keep_first(alloc_zero(&blk, DOGM_PARTICLE_BLOCK_BYTES(N)));
    particle_set_assign(h->pa, blk, h->N);
    keep_first(alloc_zero((void**)&h->rec, (N ? N : 1) * sizeof(PRec)));
    keep_first(alloc_zero((void**)&h->key0, (N ? N : 1) * sizeof(int)));
    keep_first(alloc_zero((void**)&h->pairs[0], (N ? N : 1) * sizeof(int2)));
    keep_first(alloc_zero((void**)&h->pairs[1], (N ? N : 1) * sizeof(int2)));
    h->spair = h->pairs[0];
    keep_first(alloc_zero((void**)&h->sw, (N ? N : 1) * sizeof(float)));
    h->pa_current = true;
    h->rec_valid = false;
    h->sorted_valid = false;
    keep_first(alloc_zero(&blk, DOGM_PARTICLE_BLOCK_BYTES(B)));
    particle_set_assign(h->birth, blk, h->B);
    keep_first(alloc_zero((void**)&h->grid, C * sizeof(dogm_grid_cell)));
    keep_first(alloc_zero((void**)&h->meas, C * sizeof(dogm_meas_cell)));
    keep_first(alloc_zero((void**)&h->weight_array, N * sizeof(float)));
    keep_first(alloc_zero((void**)&h->born_masses, C * sizeof(float)));
    keep_first(alloc_zero((void**)&h->free_cur, C * sizeof(float)));
    keep_first(alloc_zero((void**)&h->free_next, C * sizeof(float)));
    keep_first(alloc_zero((void**)&h->cell_start, C * sizeof(int)));
    keep_first(alloc_zero((void**)&h->cell_end, C * sizeof(int)));
    keep_first(alloc_zero((void**)&h->cell_sums, C * sizeof(CellSums)));
    keep_first(alloc_zero((void**)&h->cell_coef, C * sizeof(float4)));
    keep_first(alloc_zero((void**)&h->cell_prefix, C * sizeof(double)));
    keep_first(alloc_zero((void**)&h->blk_sum, (size_t)h->n_cell_blocks * sizeof(double)));
    keep_first(alloc_zero((void**)&h->blk_off, (size_t)h->n_cell_blocks * sizeof(double)));
    keep_first(alloc_zero((void**)&h->blk_age[0], (size_t)h->n_cell_blocks));
    keep_first(alloc_zero((void**)&h->blk_age[1], (size_t)h->n_cell_blocks));
    {
        // the shortcut pays where the grid is much larger than what the sensor sees; below ~4e6 cells the cell kernel is bound
        // by latency, not by its stores, and the bookkeeping costs more than the skipped blocks save (measured at 1.44e6 cells).
        // DOGM_B200_NO_QUIET = 1 / 0 forces it off / on (A/B runs, tests)
        const char* nq = getenv("DOGM_B200_NO_QUIET");
        h->quiet_off = nq ? nq[0] == '1' : h->C < 4000000;
    }
    keep_first(alloc_zero((void**)&h->grp_word, (size_t)h->n_blk_groups * sizeof(double)));
    keep_first(alloc_zero((void**)&h->grp_zero, ((size_t)h->n_blk_groups + 1) * sizeof(double)));
    for (int p = 0; p < kMaxPasses; p++)
    {
        h->hist[p] = nullptr;
        h->bin_base[p] = nullptr;
        h->scan_epoch[p] = 0;
    }
    // bucket sort: when the grouping table [tiles][buckets] stays small (no bands: their particle counts change per cycle)
    h->hist0_bucket = false;
    {
        const char* sm = getenv("DOGM_B200_SORT");
        h->sort_fuse_off = sm && !strcmp(sm, "radix");
    }
    h->bucket.enabled = false;
    h->bucket.samples_valid = false;
    h->bucket.bins = 0;
    h->bucket.smp_raw = nullptr;
    h->bucket.bkt = nullptr;
    h->bucket.org = nullptr;
    h->bucket.plan_key = h->bucket.plan_slot = h->bucket.plan_base = h->bucket.plan_pos = nullptr;
    {
        // one splitter per 2048 slots while that gives at most 1024 of them (buckets of ~2048 pairs: after the merge of the
        // population's two sorted runs hardly any bucket exceeds the 4096 pairs the one-pass path of k_bucket_sort takes)
        h->bucket.stride_shift = h->N <= 1024 * 2048 ? 11 : 12;
        h->bucket.n_spl = (int)(((long long)h->N + (1ll << h->bucket.stride_shift) - 1) >> h->bucket.stride_shift);
        if (h->bucket.n_spl < 1)
            h->bucket.n_spl = 1;
        const long long nb = (long long)h->bucket.n_spl + ((long long)h->C + 2047) / 2048 + 1;
        const long long bins = (nb + 31) / 32 * 32;
        // Off by default: measured at the headline size the three steps (plan 4 us, bucket numbers in the prediction +17 us,
        // grouping pass 15 us, per-bucket sort ~50 us) lose against the two radix passes (50 us) - see DESIGN.md, negative results.
        // DOGM_B200_SORT=bucket switches it on (tests run one parity case through it).
        const char* mode = getenv("DOGM_B200_SORT");
        const bool want = mode && !strcmp(mode, "bucket");
        if (want && !band && h->N > 0 && h->tiles <= 1024 && bins <= (1 << kMaxDigitBits))
        {
            h->bucket.enabled = true;
            h->bucket.bins = (int)bins;
        }
    }
    for (int p = 0; p < h->passes; p++)
    {
        size_t bins = (size_t)h->digit_bins[p];
        if (p == 0 && h->bucket.enabled && (size_t)h->bucket.bins > bins)
            bins = (size_t)h->bucket.bins;
        keep_first(alloc_zero((void**)&h->hist[p], (size_t)h->tiles * bins * sizeof(uint32_t)));
        keep_first(alloc_zero((void**)&h->bin_base[p], bins * sizeof(uint32_t)));
    }
    if (h->bucket.enabled)
    {
        keep_first(alloc_zero((void**)&h->bucket.smp_raw, (size_t)h->bucket.n_spl * sizeof(int)));
        keep_first(alloc_zero((void**)&h->bucket.plan_key, (size_t)h->bucket.n_spl * sizeof(int)));
        keep_first(alloc_zero((void**)&h->bucket.plan_slot, (size_t)h->bucket.n_spl * sizeof(int)));
        keep_first(alloc_zero((void**)&h->bucket.plan_base, (size_t)h->bucket.n_spl * sizeof(int)));
        keep_first(alloc_zero((void**)&h->bucket.plan_pos, (size_t)h->bucket.n_spl * sizeof(int)));
        keep_first(alloc_zero((void**)&h->bucket.bkt, (N ? N : 1) * sizeof(uint16_t)));
        keep_first(alloc_zero((void**)&h->bucket.org, (size_t)h->bucket.bins * sizeof(int)));
    }
    keep_first(alloc_zero((void**)&h->seg_lead, (size_t)(h->n_chunks + 1) * sizeof(SegPiece)));
    keep_first(alloc_zero((void**)&h->seg_trail, (size_t)(h->n_chunks + 1) * sizeof(SegPiece)));
    keep_first(alloc_zero((void**)&h->seg_flags, (size_t)(h->n_chunks + 1) * sizeof(int)));
    keep_first(alloc_zero((void**)&h->cdf, (N + B + 2) * sizeof(double))); // (+ padding: the bulk copy of a window reads whole 16 bytes)
    keep_first(alloc_zero((void**)&h->tile_sum, ((size_t)h->n_cdf_tiles + 3) * sizeof(double)));
    keep_first(alloc_zero((void**)&h->tile_off, ((size_t)h->n_cdf_tiles + 3) * sizeof(double)));
    keep_first(alloc_zero((void**)&h->res_start, ((size_t)div_up(h->N > 0 ? h->N : 1, kBlock)) * sizeof(int)));
    if (!e)
        e = (int)cudaMemset(h->res_start, 0x7f, ((size_t)div_up(h->N > 0 ? h->N : 1, kBlock)) * sizeof(int));
    keep_first(alloc_zero((void**)&h->chain_flags, 4 * sizeof(uint32_t)));
    {
        const char* sk = getenv("DOGM_B200_SKIP");
        h->skip_mask = sk ? (unsigned)strtoul(sk, nullptr, 0) : 0u;
    }
    h->trace_buf = nullptr;
    h->chain_epoch = 0;
    h->chain_ticket_base = 0;
    h->chain_capacity = chain_blocks_per_sm() * sm_count;
    keep_first(alloc_zero((void**)&h->ancestors, N * sizeof(int)));
    keep_first(alloc_zero((void**)&h->scal, sizeof(DeviceScalars)));
    if (e)
    {
        dogm_destroy(h);
        return e;
    }
    if (band)
    {
        for (int d = 0; d < 2; d++)
        {
            keep_first(alloc_zero((void**)&h->band.send[d], (size_t)(h->band.send_cap ? h->band.send_cap : 1) * sizeof(PRec)));
            keep_first(alloc_zero((void**)&h->band.recv[d], (size_t)(h->band.send_cap ? h->band.send_cap : 1) * sizeof(PRec)));
            keep_first(alloc_zero((void**)&h->band.halo[d], (size_t)(h->band.halo_rows ? h->band.halo_rows : 1) * gs * sizeof(float)));
        }
        const size_t out_tiles = (size_t)div_up(h->band.n_cap > 0 ? h->band.n_cap : 1, kTileItems);
        keep_first(alloc_zero((void**)&h->band.mig, (size_t)(h->band.n_cap > 0 ? h->band.n_cap : 1) + 16));
        for (int d = 0; d < 2; d++)
        {
            keep_first(alloc_zero((void**)&h->band.out_cnt[d], out_tiles * sizeof(double)));
            keep_first(alloc_zero((void**)&h->band.out_off[d], out_tiles * sizeof(double)));
        }
        keep_first(alloc_zero((void**)&h->band.out_total, 2 * sizeof(double)));
        keep_first((int)cudaMallocHost((void**)&h->band.pin, 8 * sizeof(double)));
        keep_first(alloc_zero((void**)&h->band.cnt, sizeof(BandCounts)));
        keep_first(alloc_zero((void**)&h->band.mail, sizeof(BandMail)));
        keep_first((int)cudaMallocHost((void**)&h->band.cnt_host, sizeof(BandCounts)));
        if (e)
        {
            dogm_destroy(h);
            return e;
        }
    }
    e = run_init_grid(h); // DOGM::initialize, dogm.cu:95-113
    if (e == 0)
        e = (int)cudaStreamSynchronize(h->stream);
    if (e)
    {
        dogm_destroy(h);
        return e;
    }
    if (band)
        set_particle_counts(h, 0, 0); // a band starts empty: the first cycle hands it its share of the particles
    *out = h;
    return 0;
}

extern "C" int dogm_create(const dogm_params* params, dogm_handle** out)
{
    return create_impl(params, nullptr, out);
}

namespace dogm_b200
{
// band mode: the counts of this cycle and everything derived from them (the buffers are sized for the capacities)
void set_particle_counts(dogm_handle* h, int n, int b)
{
    h->N = n;
    h->B = b;
    h->tiles = div_up(n > 0 ? n : 1, kTileItems);
    h->n_chunks = div_up(n > 0 ? n : 1, kSegChunk);
    h->n_cdf_tiles = div_up((long long)n + b > 0 ? (long long)n + b : 1, kCdfTile);
}
} // namespace dogm_b200

extern "C" void dogm_destroy(dogm_handle* h)
{
    if (!h)
        return;
    cudaStreamSynchronize(h->stream);
    cudaFree(h->pa.block);
    cudaFree(h->rec);
    cudaFree(h->key0);
    cudaFree(h->pairs[0]);
    cudaFree(h->pairs[1]);
    cudaFree(h->sw);
    cudaFree(h->birth.block);
    cudaFree(h->grid);
    if (h->grid_alt)
    {
        cudaStreamSynchronize(h->copy_stream);
        cudaFree(h->grid_alt);
        cudaEventDestroy(h->cycle_done_ev);
        cudaEventDestroy(h->grid_copy_ev[0]);
        cudaEventDestroy(h->grid_copy_ev[1]);
        cudaStreamDestroy(h->copy_stream);
    }
    cudaFree(h->meas);
    cudaFree(h->weight_array);
    cudaFree(h->born_masses);
    cudaFree(h->free_cur);
    cudaFree(h->free_next);
    cudaFree(h->cell_start);
    cudaFree(h->cell_end);
    cudaFree(h->cell_sums);
    cudaFree(h->cell_coef);
    cudaFree(h->cell_prefix);
    cudaFree(h->blk_sum);
    cudaFree(h->blk_off);
    cudaFree(h->blk_age[0]);
    cudaFree(h->blk_age[1]);
    cudaFree(h->grp_word);
    cudaFree(h->grp_zero);
    for (int p = 0; p < kMaxPasses; p++)
    {
        cudaFree(h->hist[p]);
        cudaFree(h->bin_base[p]);
    }
    cudaFree(h->bucket.smp_raw);
    cudaFree(h->bucket.plan_key);
    cudaFree(h->bucket.plan_slot);
    cudaFree(h->bucket.plan_base);
    cudaFree(h->bucket.plan_pos);
    cudaFree(h->bucket.bkt);
    cudaFree(h->bucket.org);
    cudaFree(h->seg_lead);
    cudaFree(h->seg_trail);
    cudaFree(h->seg_flags);
    cudaFree(h->cdf);
    cudaFree(h->tile_sum);
    cudaFree(h->tile_off);
    cudaFree(h->chain_flags);
    cudaFree(h->res_start);
    for (int d = 0; d < 2; d++)
    {
        cudaFree(h->band.send[d]);
        cudaFree(h->band.recv[d]);
        cudaFree(h->band.halo[d]);
    }
    cudaFree(h->band.mig);
    for (int d = 0; d < 2; d++)
    {
        cudaFree(h->band.out_cnt[d]);
        cudaFree(h->band.out_off[d]);
    }
    cudaFree(h->band.out_total);
    if (h->band.pin)
        cudaFreeHost(h->band.pin);
    cudaFree(h->band.cnt);
    for (int k = 0; k < 8; k++)
        if (h->band.stage_ev[k])
            cudaEventDestroy(h->band.stage_ev[k]);
    cudaFree(h->band.mail);
    if (h->band.cnt_host)
        cudaFreeHost(h->band.cnt_host);
    if (h->trace_buf)
    {
        trace_bind_particles(nullptr);
        trace_bind_cells(nullptr);
        trace_bind_meas(nullptr);
        cudaFree(h->trace_buf);
    }
    cudaFree(h->ancestors);
    cudaFree(h->scal);
    cudaFree(h->predict_noise);
    cudaFree(h->birth_noise);
    cudaFree(h->init_velocity);
    cudaFree(h->resample_u);
    cudaFree(h->dyn_buf);
    cudaFree(h->dyn_count);
    if (h->dyn_count_host)
        cudaFreeHost(h->dyn_count_host);
    if (h->dyn_pub_host)
        cudaFreeHost(h->dyn_pub_host);
    if (h->dyn_mapped_host)
        cudaFreeHost(h->dyn_mapped_host);
    for (auto& t : h->timed)
    {
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    for (auto& ev : h->event_pool)
        cudaEventDestroy(ev);
    if (h->timer_ready)
    {
        cudaEventDestroy(h->timer_e0);
        cudaEventDestroy(h->timer_e1);
    }
    cudaStreamDestroy(h->stream);
    delete h;
}

// ---------------------------------------------------------------------------------------------------------
// the cycle
// ---------------------------------------------------------------------------------------------------------
static int noise_ready(const dogm_handle* h, bool first_cycle)
{
    if (h->opts.noise_mode != DOGM_NOISE_INJECTED)
    {
        if (h->opts.resample_mode == DOGM_RESAMPLE_INJECTED && !h->resample_u)
            return DOGM_ERR_NOT_INITIALIZED;
        return 0;
    }
    if ((h->N > 0 && !h->predict_noise) || (h->B > 0 && !h->birth_noise) || (h->N > 0 && !h->resample_u))
        return DOGM_ERR_NOT_INITIALIZED;
    if (first_cycle && h->N > 0 && !h->init_velocity)
        return DOGM_ERR_NOT_INITIALIZED;
    return 0;
}

// updatePose, dogm.cu:161-203: only the bookkeeping happens here; the particle shift is folded into the
// prediction kernel and the grid shift into the cell kernel
static void update_pose(dogm_handle* h, float new_x, float new_y, float new_yaw)
{
    // A shift that neither the prediction nor the cell kernel has consumed yet (the stage API allows two pose updates in a
    // row) is carried over: the reference moves grid and particles inside updatePose, so two calls add up there too.
    const bool carry = h->shift.active && h->shift_particles_pending && h->shift_grid_pending;
    const int carry_x = carry ? h->shift.x_move : 0, carry_y = carry ? h->shift.y_move : 0;
    h->shift.active = carry ? 1 : 0;
    h->shift.x_move = carry_x;
    h->shift.y_move = carry_y;
    h->shift_particles_pending = h->shift_grid_pending = carry;
    if (!h->first_pose_received)
    {
        h->position_x = new_x;
        h->position_y = new_y;
        h->yaw = new_yaw;
        h->first_pose_received = true;
        return;
    }
    const float x_diff = new_x - h->position_x;
    const float y_diff = new_y - h->position_y;
    if (fabsf(x_diff) > h->params.resolution || fabsf(y_diff) > h->params.resolution)
    {
        h->shift.x_move = carry_x - (int)(x_diff / h->params.resolution);
        h->shift.y_move = carry_y - (int)(y_diff / h->params.resolution);
        h->shift.active = 1;
        h->shift_particles_pending = h->shift_grid_pending = true;
        h->position_x = new_x;
        h->position_y = new_y;
        h->yaw = new_yaw;
    }
}

static int update_grid_impl(dogm_handle* h, const dogm_meas_cell* meas, float new_x, float new_y, float new_yaw, float dt,
                            int on_device, bool sync)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    int e = noise_ready(h, !h->first_measurement_received);
    if (e)
        return e;
    // updateMeasurementGrid, dogm.cu:205-215
    h->meas_src = nullptr;
    if (meas && meas != h->meas)
    {
        h->lazy_meas.pending = false; // superseded by the caller's grid
        if (on_device && h->first_measurement_received)
        {
            // the cell kernel reads the caller's buffer and writes the handle's copy on the way: the reference's
            // separate 16*C-byte device-to-device copy (dogm.cu:207-208) costs no extra pass
            h->meas_src = meas;
        }
        else
        {
            LaunchScope ls(h, K_MEMSET, 32.0 * h->C);
            DOGM_CHECK((cudaError_t)copy_in(h->meas, meas, (size_t)h->C * sizeof(dogm_meas_cell), on_device, h->stream));
        }
    }
    if (!h->first_measurement_received)
    {
        if ((e = materialize_meas(h)))
            return e;
        e = run_init_particles(h);
        if (e)
            return e;
        h->first_measurement_received = true;
    }
    update_pose(h, new_x, new_y, new_yaw);

    if ((e = run_predict(h, dt)))
        return e;
    if ((e = run_assignment(h)))
        return e;
    if ((e = run_occupancy_update(h, dt)))
        return e;
    if ((e = run_persistent_weights(h, true)))
        return e;
    if ((e = run_birth(h)))
        return e;
    if ((e = run_resampling(h)))
        return e;
    h->cycle++;
    if (sync)
    {
        DOGM_CHECK(cudaStreamSynchronize(h->stream));
        h->cell_kernel_done = true;
    }
    return 0;
}

extern "C" int dogm_update_grid(dogm_handle* h, const dogm_meas_cell* measurement_grid, float new_x, float new_y,
                                float new_yaw, float dt, int on_device)
{
    return update_grid_impl(h, measurement_grid, new_x, new_y, new_yaw, dt, on_device, true);
}

extern "C" int dogm_update_grid_async(dogm_handle* h, const dogm_meas_cell* measurement_grid, float new_x, float new_y,
                                      float new_yaw, float dt, int on_device)
{
    return update_grid_impl(h, measurement_grid, new_x, new_y, new_yaw, dt, on_device, false);
}

extern "C" int dogm_synchronize(dogm_handle* h)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    h->cell_kernel_done = true;
    return 0;
}

#define STAGE_PROLOGUE()                                                                                               \
    if (!h)                                                                                                            \
        return DOGM_ERR_INVALID_ARGUMENT;                                                                              \
    int e = materialize_meas(h);                                                                                       \
    if (e)                                                                                                             \
        return e;
#define STAGE_EPILOGUE()                                                                                               \
    if (e)                                                                                                             \
        return e;                                                                                                      \
    DOGM_CHECK(cudaStreamSynchronize(h->stream));                                                                      \
    return 0;

extern "C" int dogm_update_measurement_grid(dogm_handle* h, const dogm_meas_cell* measurement_grid, int on_device)
{
    STAGE_PROLOGUE();
    if (h->opts.noise_mode == DOGM_NOISE_INJECTED && !h->first_measurement_received && h->N > 0 && !h->init_velocity)
        return DOGM_ERR_NOT_INITIALIZED;
    if (measurement_grid && measurement_grid != h->meas)
        DOGM_CHECK((cudaError_t)copy_in(h->meas, measurement_grid, (size_t)h->C * sizeof(dogm_meas_cell), on_device, h->stream));
    if (!h->first_measurement_received)
    {
        e = run_init_particles(h);
        h->first_measurement_received = true;
    }
    STAGE_EPILOGUE();
}

extern "C" int dogm_update_pose(dogm_handle* h, float new_x, float new_y, float new_yaw)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    update_pose(h, new_x, new_y, new_yaw);
    return 0;
}

extern "C" int dogm_initialize_particles(dogm_handle* h)
{
    STAGE_PROLOGUE();
    if (h->opts.noise_mode == DOGM_NOISE_INJECTED && h->N > 0 && !h->init_velocity)
        return DOGM_ERR_NOT_INITIALIZED;
    e = run_init_particles(h);
    h->first_measurement_received = true;
    STAGE_EPILOGUE();
}

extern "C" int dogm_particle_prediction(dogm_handle* h, float dt)
{
    STAGE_PROLOGUE();
    if (h->opts.noise_mode == DOGM_NOISE_INJECTED && h->N > 0 && !h->predict_noise)
        return DOGM_ERR_NOT_INITIALIZED;
    e = run_predict(h, dt);
    STAGE_EPILOGUE();
}

extern "C" int dogm_particle_assignment(dogm_handle* h)
{
    STAGE_PROLOGUE();
    e = run_assignment(h);
    STAGE_EPILOGUE();
}

extern "C" int dogm_grid_cell_occupancy_update(dogm_handle* h, float dt)
{
    STAGE_PROLOGUE();
    e = run_occupancy_update(h, dt);
    STAGE_EPILOGUE();
}

extern "C" int dogm_update_persistent_particles(dogm_handle* h)
{
    STAGE_PROLOGUE();
    e = run_persistent_weights(h, false);
    STAGE_EPILOGUE();
}

extern "C" int dogm_initialize_new_particles(dogm_handle* h)
{
    STAGE_PROLOGUE();
    if (h->opts.noise_mode == DOGM_NOISE_INJECTED && h->B > 0 && !h->birth_noise)
        return DOGM_ERR_NOT_INITIALIZED;
    e = run_birth(h);
    STAGE_EPILOGUE();
}

extern "C" int dogm_statistical_moments(dogm_handle* h)
{
    // the velocity moments are produced by the fused cell kernel of dogm_grid_cell_occupancy_update
    STAGE_PROLOGUE();
    STAGE_EPILOGUE();
}

extern "C" int dogm_resampling(dogm_handle* h)
{
    STAGE_PROLOGUE();
    if (h->N > 0 && !h->resample_u &&
        (h->opts.noise_mode == DOGM_NOISE_INJECTED || h->opts.resample_mode == DOGM_RESAMPLE_INJECTED))
        return DOGM_ERR_NOT_INITIALIZED;
    e = run_resampling(h);
    h->cycle++;
    STAGE_EPILOGUE();
}

// ---------------------------------------------------------------------------------------------------------
// read-out
// ---------------------------------------------------------------------------------------------------------
static int copy_out(dogm_handle* h, void* dst_host, const void* src_device, size_t bytes)
{
    if (!h || !dst_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (bytes == 0)
        return 0;
    DOGM_CHECK(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, h->stream));
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_get_grid_cells(dogm_handle* h, dogm_grid_cell* out_host)
{
    return copy_out(h, out_host, h ? h->grid : nullptr, h ? (size_t)h->C * sizeof(dogm_grid_cell) : 0);
}
// Pipelined getGridCells: the copy is enqueued on a stream of its own behind everything the handle has enqueued so far and
// overlaps whatever is enqueued next (the next cycle's measurement upload and kernels); from then on the handle owns two
// GridCell buffers and the cell kernel alternates between them.
extern "C" int dogm_get_grid_cells_begin(dogm_handle* h, dogm_grid_cell* out_host)
{
    if (!h || !out_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (h->band.enabled)
        return DOGM_ERR_UNSUPPORTED;
    DOGM_CHECK(cudaSetDevice(h->device));
    if (!h->grid_alt)
    {
        DOGM_CHECK(cudaMalloc((void**)&h->grid_alt, (size_t)h->C * sizeof(dogm_grid_cell)));
        DOGM_CHECK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        DOGM_CHECK(cudaEventCreateWithFlags(&h->cycle_done_ev, cudaEventDisableTiming));
        DOGM_CHECK(cudaEventCreateWithFlags(&h->grid_copy_ev[0], cudaEventDisableTiming));
        DOGM_CHECK(cudaEventCreateWithFlags(&h->grid_copy_ev[1], cudaEventDisableTiming));
        // (a caller may read cells of the other buffer through the stage API before the first full cycle has written it)
        DOGM_CHECK(cudaMemcpyAsync(h->grid_alt, h->grid, (size_t)h->C * sizeof(dogm_grid_cell), cudaMemcpyDeviceToDevice, h->stream));
    }
    DOGM_CHECK(cudaEventRecord(h->cycle_done_ev, h->stream));
    DOGM_CHECK(cudaStreamWaitEvent(h->copy_stream, h->cycle_done_ev, 0));
    DOGM_CHECK(cudaMemcpyAsync(out_host, h->grid, (size_t)h->C * sizeof(dogm_grid_cell), cudaMemcpyDeviceToHost, h->copy_stream));
    DOGM_CHECK(cudaEventRecord(h->grid_copy_ev[0], h->copy_stream));
    h->grid_copy_busy[0] = true;
    h->grid_swap_pending = true;
    return 0;
}

extern "C" int dogm_get_grid_cells_wait(dogm_handle* h)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (!h->grid_alt)
        return 0;
    DOGM_CHECK(cudaSetDevice(h->device));
    DOGM_CHECK(cudaStreamSynchronize(h->copy_stream));
    return 0;
}

extern "C" int dogm_get_measurement_cells(dogm_handle* h, dogm_meas_cell* out_host)
{
    if (h && materialize_meas(h))
        return (int)cudaGetLastError();
    return copy_out(h, out_host, h ? h->meas : nullptr, h ? (size_t)h->C * sizeof(dogm_meas_cell) : 0);
}
extern "C" int dogm_get_particles(dogm_handle* h, void* out_block)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    int e = ensure_soa(h); // between prediction and resampling the particles live in the record buffers
    if (e)
        return e;
    return copy_out(h, out_block, h->pa.block, DOGM_PARTICLE_BLOCK_BYTES(h->N));
}
extern "C" int dogm_get_birth_particles(dogm_handle* h, void* out_block)
{
    return copy_out(h, out_block, h ? h->birth.block : nullptr, h ? DOGM_PARTICLE_BLOCK_BYTES(h->B) : 0);
}
extern "C" int dogm_get_weight_array(dogm_handle* h, float* out_host)
{
    return copy_out(h, out_host, h ? h->weight_array : nullptr, h ? (size_t)h->N * sizeof(float) : 0);
}
extern "C" int dogm_get_born_masses(dogm_handle* h, float* out_host)
{
    return copy_out(h, out_host, h ? h->born_masses : nullptr, h ? (size_t)h->C * sizeof(float) : 0);
}
extern "C" int dogm_get_resampled_indices(dogm_handle* h, int* out_host)
{
    return copy_out(h, out_host, h ? h->ancestors : nullptr, h ? (size_t)h->N * sizeof(int) : 0);
}
extern "C" int dogm_get_joint_weight_accum(dogm_handle* h, double* out_host)
{
    return copy_out(h, out_host, h ? h->cdf : nullptr, h ? ((size_t)h->N + h->B) * sizeof(double) : 0);
}
extern "C" int dogm_get_cell_ranges(dogm_handle* h, int* start_host, int* end_host)
{
    if (!h || !start_host || !end_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (h->ranges_in_soa)
    { // between particleAssignment and gridCellOccupancyUpdate the ranges live in the compact arrays
        int e = copy_out(h, start_host, h->cell_start, (size_t)h->C * sizeof(int));
        if (e)
            return e;
        return copy_out(h, end_host, h->cell_end, (size_t)h->C * sizeof(int));
    }
    // afterwards they are GridCell.start_idx / end_idx (the cell kernel consumed and reset the compact arrays)
    DOGM_CHECK(cudaMemcpy2DAsync(start_host, sizeof(int), &h->grid[0].start_idx, sizeof(dogm_grid_cell), sizeof(int),
                                 (size_t)h->C, cudaMemcpyDeviceToHost, h->stream));
    DOGM_CHECK(cudaMemcpy2DAsync(end_host, sizeof(int), &h->grid[0].end_idx, sizeof(dogm_grid_cell), sizeof(int),
                                 (size_t)h->C, cudaMemcpyDeviceToHost, h->stream));
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_get_grid_size(const dogm_handle* h) { return h ? h->gs : 0; }
extern "C" int dogm_get_grid_cell_count(const dogm_handle* h) { return h ? h->C : 0; }
extern "C" int dogm_get_particle_count(const dogm_handle* h) { return h ? h->N : 0; }
extern "C" int dogm_get_new_born_particle_count(const dogm_handle* h) { return h ? h->B : 0; }
extern "C" float dogm_get_resolution(const dogm_handle* h) { return h ? h->params.resolution : 0.0f; }
extern "C" float dogm_get_position_x(const dogm_handle* h) { return h ? h->position_x : 0.0f; }
extern "C" float dogm_get_position_y(const dogm_handle* h) { return h ? h->position_y : 0.0f; }
extern "C" float dogm_get_yaw(const dogm_handle* h) { return h ? h->yaw : 0.0f; }
extern "C" uint32_t dogm_get_cycle_counter(const dogm_handle* h) { return h ? h->cycle : 0u; }

extern "C" int dogm_get_device_ptrs(dogm_handle* h, dogm_device_ptrs* out)
{
    if (!h || !out)
        return DOGM_ERR_INVALID_ARGUMENT;
    out->grid_cell_array = h->grid;
    materialize_meas(h);
    ensure_soa(h);
    out->particle_array = h->pa.block;
    out->particle_array_next = h->pa.block; // resampling writes the next population in place (no 28*N-byte publish copy)
    out->birth_particle_array = h->birth.block;
    out->meas_cell_array = h->meas;
    out->weight_array = h->weight_array;
    out->born_masses_array = h->born_masses;
    out->resampled_idx_array = h->ancestors;
    out->joint_weight_accum = h->cdf;
    out->cell_start_array = h->cell_start;
    out->cell_end_array = h->cell_end;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// options and parity hooks
// ---------------------------------------------------------------------------------------------------------
extern "C" int dogm_set_options(dogm_handle* h, const dogm_options* opts)
{
    if (!h || !opts)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (opts->resample_mode < DOGM_RESAMPLE_SYSTEMATIC || opts->resample_mode > DOGM_RESAMPLE_INJECTED)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (opts->noise_mode != DOGM_NOISE_PHILOX && opts->noise_mode != DOGM_NOISE_INJECTED)
        return DOGM_ERR_INVALID_ARGUMENT;
    h->opts = *opts;
    return 0;
}

extern "C" int dogm_get_options(const dogm_handle* h, dogm_options* out)
{
    if (!h || !out)
        return DOGM_ERR_INVALID_ARGUMENT;
    *out = h->opts;
    return 0;
}

static int set_buffer(dogm_handle* h, void** slot, const void* src, size_t bytes, int on_device)
{
    if (!src || bytes == 0)
        return 0;
    if (!*slot)
        DOGM_CHECK(cudaMalloc(slot, bytes));
    DOGM_CHECK((cudaError_t)copy_in(*slot, src, bytes, on_device, h->stream));
    return 0;
}

extern "C" int dogm_set_noise(dogm_handle* h, const float* predict_noise, const float* birth_noise,
                              const float* init_velocity, const float* resample_unit, int on_device)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    int e = 0;
    e |= set_buffer(h, (void**)&h->predict_noise, predict_noise, (size_t)h->N * sizeof(float4), on_device);
    e |= set_buffer(h, (void**)&h->birth_noise, birth_noise, (size_t)h->B * sizeof(float2), on_device);
    e |= set_buffer(h, (void**)&h->init_velocity, init_velocity, (size_t)h->N * sizeof(float2), on_device);
    e |= set_buffer(h, (void**)&h->resample_u, resample_unit, (size_t)h->N * sizeof(float), on_device);
    if (e)
        return e;
    DOGM_CHECK(cudaStreamSynchronize(h->stream)); // the caller may reuse its buffers
    return 0;
}

extern "C" int dogm_export_philox_noise(dogm_handle* h, uint32_t cycle, float* predict_noise, float* birth_noise,
                                        float* init_velocity, float* resample_unit)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    float4* d_p = nullptr;
    float2 *d_b = nullptr, *d_i = nullptr;
    float* d_r = nullptr;
    const size_t N = (size_t)h->N, B = (size_t)h->B;
    if (predict_noise && N)
        DOGM_CHECK(cudaMalloc(&d_p, N * sizeof(float4)));
    if (birth_noise && B)
        DOGM_CHECK(cudaMalloc(&d_b, B * sizeof(float2)));
    if (init_velocity && N)
        DOGM_CHECK(cudaMalloc(&d_i, N * sizeof(float2)));
    if (resample_unit && N)
        DOGM_CHECK(cudaMalloc(&d_r, N * sizeof(float)));
    int e = run_export_noise(h, cycle, d_p, d_b, d_i, d_r);
    if (!e && d_p)
        e = (int)cudaMemcpyAsync(predict_noise, d_p, N * sizeof(float4), cudaMemcpyDeviceToHost, h->stream);
    if (!e && d_b)
        e = (int)cudaMemcpyAsync(birth_noise, d_b, B * sizeof(float2), cudaMemcpyDeviceToHost, h->stream);
    if (!e && d_i)
        e = (int)cudaMemcpyAsync(init_velocity, d_i, N * sizeof(float2), cudaMemcpyDeviceToHost, h->stream);
    if (!e && d_r)
        e = (int)cudaMemcpyAsync(resample_unit, d_r, N * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    if (!e)
        e = (int)cudaStreamSynchronize(h->stream);
    cudaFree(d_p);
    cudaFree(d_b);
    cudaFree(d_i);
    cudaFree(d_r);
    return e;
}

extern "C" int dogm_set_particles(dogm_handle* h, const void* block, int on_device)
{
    if (!h || !block)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK((cudaError_t)copy_in(h->pa.block, block, DOGM_PARTICLE_BLOCK_BYTES(h->N), on_device, h->stream));
    h->hist0_valid = false;
    h->bucket.samples_valid = false;
    h->pa_current = true;
    h->rec_valid = false;
    h->sorted_valid = false;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_set_birth_particles(dogm_handle* h, const void* block, int on_device)
{
    if (!h || !block)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK((cudaError_t)copy_in(h->birth.block, block, DOGM_PARTICLE_BLOCK_BYTES(h->B), on_device, h->stream));
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_set_grid_cells(dogm_handle* h, const dogm_grid_cell* cells, int on_device)
{
    if (!h || !cells)
        return DOGM_ERR_INVALID_ARGUMENT;
    h->dyn_list_valid = false;
    if (h->grid_alt && h->grid_copy_busy[0])
    { // a pipelined read-out of this buffer may still be on its way: it gets the cells it was begun for
        DOGM_CHECK(cudaStreamWaitEvent(h->stream, h->grid_copy_ev[0], 0));
        h->grid_copy_busy[0] = false;
    }
    DOGM_CHECK(cudaMemsetAsync(h->blk_age[0], 0, (size_t)h->n_cell_blocks, h->stream)); // (no block is known to be empty any more)
    DOGM_CHECK((cudaError_t)copy_in(h->grid, cells, (size_t)h->C * sizeof(dogm_grid_cell), on_device, h->stream));
    int e = run_extract_free_mass(h);
    if (e)
        return e;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_set_measurement_cells(dogm_handle* h, const dogm_meas_cell* cells, int on_device)
{
    if (!h || !cells)
        return DOGM_ERR_INVALID_ARGUMENT;
    h->lazy_meas.pending = false;
    DOGM_CHECK(cudaMemsetAsync(h->blk_age[0], 0, (size_t)h->n_cell_blocks, h->stream));
    DOGM_CHECK((cudaError_t)copy_in(h->meas, cells, (size_t)h->C * sizeof(dogm_meas_cell), on_device, h->stream));
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_set_pose(dogm_handle* h, float x, float y, float yaw)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    h->position_x = x;
    h->position_y = y;
    h->yaw = yaw;
    h->first_pose_received = true;
    return 0;
}

extern "C" int dogm_search_ancestors_f32(dogm_handle* h, const float* cdf, int n_cdf, const float* sorted_draws,
                                         int n_draws, int* out)
{
    if (!h || !cdf || !sorted_draws || !out || n_cdf <= 0 || n_draws < 0)
        return DOGM_ERR_INVALID_ARGUMENT;
    float *d_cdf = nullptr, *d_draws = nullptr;
    int* d_out = nullptr;
    DOGM_CHECK(cudaMalloc(&d_cdf, (size_t)n_cdf * sizeof(float)));
    DOGM_CHECK(cudaMalloc(&d_draws, (size_t)(n_draws > 0 ? n_draws : 1) * sizeof(float)));
    DOGM_CHECK(cudaMalloc(&d_out, (size_t)(n_draws > 0 ? n_draws : 1) * sizeof(int)));
    int e = (int)cudaMemcpyAsync(d_cdf, cdf, (size_t)n_cdf * sizeof(float), cudaMemcpyHostToDevice, h->stream);
    if (!e)
        e = (int)cudaMemcpyAsync(d_draws, sorted_draws, (size_t)n_draws * sizeof(float), cudaMemcpyHostToDevice, h->stream);
    if (!e)
        e = run_search_ancestors_f32(h, d_cdf, n_cdf, d_draws, n_draws, d_out);
    if (!e)
        e = (int)cudaMemcpyAsync(out, d_out, (size_t)n_draws * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (!e)
        e = (int)cudaStreamSynchronize(h->stream);
    cudaFree(d_cdf);
    cudaFree(d_draws);
    cudaFree(d_out);
    return e;
}

static int ensure_dyn_counter(dogm_handle* h)
{
    if (!h->dyn_count)
    {
        DOGM_CHECK(cudaMalloc(&h->dyn_count, 2 * sizeof(int))); // [0] fused into the cycle, [1] stand-alone pass
        DOGM_CHECK(cudaMallocHost(&h->dyn_count_host, sizeof(int)));
        DOGM_CHECK(cudaHostAlloc((void**)&h->dyn_pub_host, 2 * sizeof(int), cudaHostAllocMapped));
        h->dyn_pub_host[0] = h->dyn_pub_host[1] = 0;
        DOGM_CHECK(cudaHostGetDevicePointer((void**)&h->dyn_pub_dev, h->dyn_pub_host, 0));
    }
    return 0;
}

extern "C" int dogm_set_dynamic_cell_filter(dogm_handle* h, float min_occupancy, float min_velocity, int capacity)
{
    if (!h || capacity < 0)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    h->dyn_filter_on = false;
    h->dyn_list_valid = false;
    if (capacity == 0)
        return 0; // filter switched off
    int e = ensure_dyn_counter(h);
    if (e)
        return e;
    if (capacity > h->dyn_filter_capacity || !h->dyn_mapped_host)
    {
        if (h->dyn_mapped_host)
            cudaFreeHost(h->dyn_mapped_host);
        h->dyn_mapped_host = nullptr;
        DOGM_CHECK(cudaHostAlloc((void**)&h->dyn_mapped_host, (size_t)capacity * sizeof(dogm_dynamic_cell), cudaHostAllocMapped));
        DOGM_CHECK(cudaHostGetDevicePointer((void**)&h->dyn_mapped_dev, h->dyn_mapped_host, 0));
    }
    DOGM_CHECK(cudaMemset(h->dyn_count, 0, 2 * sizeof(int)));
    h->dyn_filter_capacity = capacity;
    h->dyn_filter_occ = min_occupancy;
    h->dyn_filter_vel = min_velocity;
    h->dyn_filter_on = true;
    return 0;
}

extern "C" int dogm_extract_dynamic_cells(dogm_handle* h, float min_occupancy, float min_velocity,
                                          dogm_dynamic_cell* out_host, int capacity, int* out_count)
{
    if (!h || !out_count || capacity < 0 || (capacity > 0 && !out_host))
        return DOGM_ERR_INVALID_ARGUMENT;
    int e0 = ensure_dyn_counter(h);
    if (e0)
        return e0;
    // fast path: the cell kernel of the last cycle already produced the list for exactly this filter
    if (h->dyn_filter_on && h->dyn_list_valid && h->dyn_list_cycle + 1 == h->cycle && min_occupancy == h->dyn_filter_occ &&
        min_velocity == h->dyn_filter_vel)
    {
        int found = -1;
        if (h->dyn_pub_pending)
        { // wait for the publication written behind the cell kernel, not for the end of the cycle
            volatile int* pub = h->dyn_pub_host;
            for (unsigned spin = 1; found < 0; spin++)
            {
                if (pub[1] == h->dyn_pub_seq)
                {
                    std::atomic_thread_fence(std::memory_order_acquire); // the count is read after the sequence number
                    found = pub[0];
                    h->cell_kernel_done = true; // (the publisher runs behind the cell kernel)
                }
                else if ((spin & 0x3ffu) == 0)
                {
                    const cudaError_t q = cudaStreamQuery(h->stream);
                    if (q == cudaSuccess)
                    { // nothing left in flight: the word is final one way or the other
                        if (pub[1] == h->dyn_pub_seq)
                        {
                            std::atomic_thread_fence(std::memory_order_acquire);
                            found = pub[0];
                        }
                        break;
                    }
                    if (q != cudaErrorNotReady)
                        return (int)q;
                }
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire); // the list is read after the sequence number, not before
        if (found < 0)
        {
            DOGM_CHECK(cudaMemcpyAsync(h->dyn_count_host, h->dyn_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            DOGM_CHECK(cudaStreamSynchronize(h->stream));
            found = *h->dyn_count_host;
            h->cell_kernel_done = true;
        }
        if (found <= h->dyn_filter_capacity || capacity <= h->dyn_filter_capacity)
        {
            *out_count = found;
            int n_copy = found < capacity ? found : capacity;
            if (n_copy > h->dyn_filter_capacity)
                n_copy = h->dyn_filter_capacity;
            if (n_copy > 0)
                memcpy(out_host, h->dyn_mapped_host, (size_t)n_copy * sizeof(dogm_dynamic_cell));
            return 0;
        }
    }
    if (capacity > h->dyn_capacity)
    {
        cudaFree(h->dyn_buf);
        h->dyn_buf = nullptr;
        h->dyn_capacity = 0;
        DOGM_CHECK(cudaMalloc(&h->dyn_buf, (size_t)capacity * sizeof(dogm_dynamic_cell)));
        h->dyn_capacity = capacity;
    }
    int e = run_extract_dynamic_cells(h, min_occupancy, min_velocity, h->dyn_buf, capacity, h->dyn_count + 1);
    if (e)
        return e;
    DOGM_CHECK(cudaMemcpyAsync(h->dyn_count_host, h->dyn_count + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    const int found = *h->dyn_count_host;
    *out_count = found;
    const int n_copy = found < capacity ? found : capacity;
    if (n_copy > 0)
    {
        DOGM_CHECK(cudaMemcpyAsync(out_host, h->dyn_buf, (size_t)n_copy * sizeof(dogm_dynamic_cell), cudaMemcpyDeviceToHost,
                                   h->stream));
        DOGM_CHECK(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// band mode (include/dogm_b200.h, "Band mode"): the cycle in phases, the orchestrator moves data between the bands
// ---------------------------------------------------------------------------------------------------------
// band phases end with a scalar for the host: copied into pinned memory on the band's stream, one synchronisation
static int band_sync_and_fetch(dogm_handle* h, void* dst_host, const void* src_device, size_t bytes)
{
    if (dst_host && bytes)
        DOGM_CHECK(cudaMemcpyAsync(h->band.pin, src_device, bytes, cudaMemcpyDeviceToHost, h->stream));
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    if (dst_host && bytes)
        memcpy(dst_host, h->band.pin, bytes);
    return 0;
}

extern "C" int dogm_create_band(const dogm_params* params, const dogm_band_config* band, dogm_handle** out)
{
    if (!band)
        return DOGM_ERR_INVALID_ARGUMENT;
    return create_impl(params, band, out);
}

#define BAND_PROLOGUE()                                                                                                \
    if (!h || !h->band.enabled)                                                                                        \
        return DOGM_ERR_INVALID_ARGUMENT;                                                                              \
    if (h->opts.noise_mode != DOGM_NOISE_PHILOX || h->opts.resample_mode == DOGM_RESAMPLE_INJECTED)                    \
        return DOGM_ERR_UNSUPPORTED;                                                                                   \
    int e = 0;

extern "C" void* dogm_band_buffer(dogm_handle* h, int which)
{
    if (!h || !h->band.enabled)
        return nullptr;
    switch (which)
    {
    case DOGM_BAND_SEND_LO: return h->band.send[0];
    case DOGM_BAND_SEND_HI: return h->band.send[1];
    case DOGM_BAND_RECV_LO: return h->band.recv[0];
    case DOGM_BAND_RECV_HI: return h->band.recv[1];
    case DOGM_BAND_HALO_LO: return h->band.halo[0];
    case DOGM_BAND_HALO_HI: return h->band.halo[1];
    case DOGM_BAND_EDGE_LO: return h->free_cur;
    case DOGM_BAND_EDGE_HI: return h->free_cur + (size_t)(h->band.rows - h->band.halo_rows) * h->gs;
    default: return nullptr;
    }
}

extern "C" int dogm_band_counts(dogm_handle* h, int* particles, int* birth_particles)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (particles)
        *particles = h->N;
    if (birth_particles)
        *birth_particles = h->B;
    return 0;
}

// Pure host arithmetic of the band mode, exported so that an orchestrator (and a CPU test) can check that the parts of
// consecutive bands tile the slot ranges without gap or overlap.
extern "C" int dogm_band_slot_range(double mass_before, double mass_local, double mass_total, int slots_total, int* first,
                                    int* past)
{
    if (!first || !past)
        return DOGM_ERR_INVALID_ARGUMENT;
    band_slot_range(mass_before, mass_local, mass_total, slots_total, first, past);
    return 0;
}

extern "C" int dogm_band_output_range(uint64_t seed, uint32_t cycle, int resample_mode, long long n_glob, double weight_before,
                                      double weight_local, double weight_total, long long* first, long long* past)
{
    if (!first || !past)
        return DOGM_ERR_INVALID_ARGUMENT;
    long long i_lo = 0, i_hi = 0;
    band_output_range(seed, cycle, resample_mode == DOGM_RESAMPLE_SYSTEMATIC, n_glob, weight_before, weight_local, weight_total, &i_lo,
                      &i_hi);
    *first = i_lo;
    *past = i_hi;
    return 0;
}

extern "C" int dogm_band_init_masses(dogm_handle* h, const dogm_meas_cell* measurement_band, int on_device, double* mass_local)
{
    BAND_PROLOGUE();
    if (!mass_local)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (measurement_band && measurement_band != h->meas)
        DOGM_CHECK((cudaError_t)copy_in(h->meas, measurement_band, (size_t)h->C * sizeof(dogm_meas_cell), on_device, h->stream));
    e = run_init_masses(h);
    if (e)
        return e;
    if ((e = band_sync_and_fetch(h, &h->band.born_local, &h->scal->born_total, sizeof(double))))
        return e;
    *mass_local = h->band.born_local;
    return 0;
}

extern "C" int dogm_band_init_particles(dogm_handle* h, double mass_before, double mass_total, int* particles)
{
    BAND_PROLOGUE();
    int first, past;
    band_slot_range(mass_before, h->band.born_local, mass_total, h->band.n_glob, &first, &past);
    int n = past - first;
    if (n > h->band.n_cap)
        return DOGM_ERR_INVALID_ARGUMENT; // the band's share does not fit its capacity
    h->band.born_base = mass_before;
    h->band.birth_slot_base = first;
    h->band.pin[4] = mass_total; // (pinned: the asynchronous copy reads it when the stream gets there)
    DOGM_CHECK(cudaMemcpyAsync(&h->scal->born_total, &h->band.pin[4], sizeof(double), cudaMemcpyHostToDevice, h->stream));
    set_particle_counts(h, n, 0);
    e = run_init_fill(h);
    h->first_measurement_received = true;
    if (e)
        return e;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    if (particles)
        *particles = n;
    return 0;
}

extern "C" int dogm_band_predict(dogm_handle* h, float new_x, float new_y, float new_yaw, float dt, int* send_lo, int* send_hi)
{
    BAND_PROLOGUE();
    update_pose(h, new_x, new_y, new_yaw);
    e = run_predict(h, dt);
    if (e)
        return e;
    h->shift_particles_pending = false;
    int counts[2] = {0, 0};
    if (h->N > 0)
    {
        if ((e = run_band_outbox(h)))
            return e;
        double totals[2] = {0.0, 0.0};
        if ((e = band_sync_and_fetch(h, totals, h->band.out_total, sizeof(totals))))
            return e;
        counts[0] = (int)totals[0];
        counts[1] = (int)totals[1];
    }
    else
        DOGM_CHECK(cudaStreamSynchronize(h->stream));
    if (counts[0] > h->band.send_cap || counts[1] > h->band.send_cap)
        return DOGM_ERR_INVALID_ARGUMENT; // more particles crossed an edge than the exchange boxes hold
    if (send_lo)
        *send_lo = counts[0];
    if (send_hi)
        *send_hi = counts[1];
    return 0;
}

extern "C" int dogm_band_append(dogm_handle* h, int recv_lo, int recv_hi)
{
    BAND_PROLOGUE();
    if (recv_lo < 0 || recv_hi < 0 || recv_lo > h->band.send_cap || recv_hi > h->band.send_cap)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (h->N == 0 && recv_lo + recv_hi > 0)
    { // nothing was predicted here: the records are the band's whole population
        h->pa_current = false;
        h->rec_valid = true;
        h->sorted_valid = false;
    }
    e = run_band_append(h, recv_lo, recv_hi);
    if (e)
        return e;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

// The exchange step of a band in one call: the neighbours' outboxes and edge rows are copied on the band's own stream
// (direct NVLink copies once dogm_enable_peer_access has been called for the two devices, staged through the host otherwise),
// the received particles are appended, one synchronisation at the end.
extern "C" int dogm_band_receive(dogm_handle* h, const void* outbox_of_lower_neighbour, int recv_lo,
                                 const void* outbox_of_upper_neighbour, int recv_hi, const void* edge_rows_of_lower_neighbour,
                                 const void* edge_rows_of_upper_neighbour)
{
    BAND_PROLOGUE();
    if (recv_lo < 0 || recv_hi < 0 || recv_lo > h->band.send_cap || recv_hi > h->band.send_cap ||
        (recv_lo > 0 && !outbox_of_lower_neighbour) || (recv_hi > 0 && !outbox_of_upper_neighbour))
        return DOGM_ERR_INVALID_ARGUMENT;
    if (recv_lo > 0)
        DOGM_CHECK(cudaMemcpyAsync(h->band.recv[0], outbox_of_lower_neighbour, (size_t)recv_lo * sizeof(PRec), cudaMemcpyDefault,
                                   h->stream));
    if (recv_hi > 0)
        DOGM_CHECK(cudaMemcpyAsync(h->band.recv[1], outbox_of_upper_neighbour, (size_t)recv_hi * sizeof(PRec), cudaMemcpyDefault,
                                   h->stream));
    const size_t halo_bytes = (size_t)h->band.halo_rows * h->gs * sizeof(float);
    if (halo_bytes && edge_rows_of_lower_neighbour)
        DOGM_CHECK(cudaMemcpyAsync(h->band.halo[0], edge_rows_of_lower_neighbour, halo_bytes, cudaMemcpyDefault, h->stream));
    if (halo_bytes && edge_rows_of_upper_neighbour)
        DOGM_CHECK(cudaMemcpyAsync(h->band.halo[1], edge_rows_of_upper_neighbour, halo_bytes, cudaMemcpyDefault, h->stream));
    if (h->N == 0 && recv_lo + recv_hi > 0)
    { // nothing was predicted here: the records are the band's whole population
        h->pa_current = false;
        h->rec_valid = true;
        h->sorted_valid = false;
    }
    e = run_band_append(h, recv_lo, recv_hi);
    if (e)
        return e;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_band_update(dogm_handle* h, const dogm_meas_cell* measurement_band, int on_device, float dt, int halo_valid,
                                double* born_local)
{
    BAND_PROLOGUE();
    if (!born_local)
        return DOGM_ERR_INVALID_ARGUMENT;
    h->meas_src = nullptr;
    if (measurement_band && measurement_band != h->meas)
    {
        if (on_device)
            h->meas_src = measurement_band; // the cell kernel copies it on the way
        else
            DOGM_CHECK((cudaError_t)copy_in(h->meas, measurement_band, (size_t)h->C * sizeof(dogm_meas_cell), 0, h->stream));
    }
    h->band.halo_valid = halo_valid ? 1 : 0;
    if (h->shift_grid_pending && h->shift.active && h->band.rows < h->band.G &&
        (h->shift.y_move > h->band.halo_rows || -h->shift.y_move > h->band.halo_rows))
        return DOGM_ERR_INVALID_ARGUMENT; // the shift pulls in more rows of the neighbour than the halo holds
    if ((e = run_assignment(h)))
        return e;
    if ((e = run_occupancy_update(h, dt)))
        return e;
    if ((e = run_persistent_weights(h, true)))
        return e;
    if ((e = run_born_scan(h)))
        return e;
    if ((e = band_sync_and_fetch(h, &h->band.born_local, &h->scal->born_total, sizeof(double))))
        return e;
    *born_local = h->band.born_local;
    return 0;
}

extern "C" int dogm_band_birth(dogm_handle* h, double born_before, double born_total, double* weight_local)
{
    BAND_PROLOGUE();
    if (!weight_local)
        return DOGM_ERR_INVALID_ARGUMENT;
    int first, past;
    band_slot_range(born_before, h->band.born_local, born_total, h->band.b_glob, &first, &past);
    int b = past - first;
    if (b > h->band.b_cap)
        return DOGM_ERR_INVALID_ARGUMENT;
    h->band.born_base = born_before;
    h->band.birth_slot_base = first;
    DOGM_CHECK(cudaMemcpyAsync(&h->scal->born_total, &born_total, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    set_particle_counts(h, h->N, b);
    if ((e = run_birth_fill(h)))
        return e;
    if ((e = run_cdf(h)))
        return e;
    h->band.weight_local = 0.0;
    const bool any = h->N + h->B > 0;
    if ((e = band_sync_and_fetch(h, any ? &h->band.weight_local : nullptr, &h->scal->weight_total, any ? sizeof(double) : 0)))
        return e;
    *weight_local = h->band.weight_local;
    return 0;
}

extern "C" int dogm_band_resample(dogm_handle* h, double weight_before, double weight_total, int* particles_out)
{
    BAND_PROLOGUE();
    long long i_lo = 0, i_hi = 0;
    dogm_band_output_range(h->opts.seed, h->cycle, h->opts.resample_mode, h->band.n_glob, weight_before, h->band.weight_local,
                           weight_total, &i_lo, &i_hi);
    const long long n_out = i_hi - i_lo;
    if (n_out > h->band.n_cap)
        return DOGM_ERR_INVALID_ARGUMENT;
    h->band.cdf_base = weight_before;
    h->band.out_base = i_lo;
    h->band.n_out = (int)n_out;
    DOGM_CHECK(cudaMemcpyAsync(&h->scal->weight_total, &weight_total, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    e = run_resample_gather(h);
    if (e)
        return e;
    set_particle_counts(h, (int)n_out, h->B);
    h->cycle++;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    if (particles_out)
        *particles_out = (int)n_out;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// device-paced cycle: the whole cycle of a band is enqueued in one go; the particle counts of the cycle and the bands'
// messages stay on the devices (BandCounts / BandMail, kernels_particles.cu), the host looks once, at the end
// ---------------------------------------------------------------------------------------------------------
extern "C" void* dogm_band_mailbox(dogm_handle* h)
{
    return (h && h->band.enabled) ? (void*)h->band.mail : nullptr;
}

extern "C" int dogm_band_link(dogm_handle* h, int rank, int n_bands, void* const* mailboxes)
{
    if (!h || !h->band.enabled || !mailboxes || n_bands < 1 || n_bands > kMaxBands || rank < 0 || rank >= n_bands)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (mailboxes[rank] != (void*)h->band.mail)
        return DOGM_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < kMaxBands; k++)
        h->band.peer_mail[k] = k < n_bands ? (BandMail*)mailboxes[k] : nullptr;
    h->band.group_size = n_bands;
    h->band.group_rank = rank;
    h->band.seq = 0;
    DOGM_CHECK(cudaMemsetAsync(h->band.mail, 0, sizeof(BandMail), h->stream));
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dogm_band_cycle_enqueue(dogm_handle* h, int stages, const dogm_meas_cell* measurement_band, int on_device, float new_x,
                                       float new_y, float new_yaw, float dt, const void* outbox_of_lower_neighbour,
                                       const void* outbox_of_upper_neighbour, const void* edge_rows_of_lower_neighbour,
                                       const void* edge_rows_of_upper_neighbour)
{
    BAND_PROLOGUE();
    if (h->band.group_size < 1 || !h->band.cnt)
        return DOGM_ERR_NOT_INITIALIZED; // dogm_band_link first
    if (!h->first_measurement_received)
        return DOGM_ERR_NOT_INITIALIZED; // the first cycle's initialisation runs through dogm_band_init_*
    const int R = h->band.group_size, me = h->band.group_rank;
    if ((me > 0 && !outbox_of_lower_neighbour) || (me + 1 < R && !outbox_of_upper_neighbour))
        return DOGM_ERR_INVALID_ARGUMENT;
    auto stamp = [&](int k) { // stage boundaries on the band's stream (profiling only: an event node interrupts the launch chain)
        if (h->band.profile)
            cudaEventRecord(h->band.stage_ev[k], h->stream);
    };
    if (stages & DOGM_BAND_STAGE_PREDICT)
    { // prediction, the records that leave, their counts to the neighbours
        stamp(0);
        h->band.seq++;
        update_pose(h, new_x, new_y, new_yaw);
        h->band.n_pred = h->N;
        if ((e = run_predict(h, dt)))
            return e;
        h->shift_particles_pending = false;
        if (h->N > 0 && (e = run_band_outbox(h)))
            return e;
        if ((e = run_band_publish_sent(h)))
            return e;
        stamp(1);
    }
    if (stages & DOGM_BAND_STAGE_UPDATE)
    { // the neighbours' counts and records; sort, per-cell sums, occupancy update; this band's born mass to all bands
        const int n_pred = h->band.n_pred;
        if ((e = run_band_collect_sent(h, n_pred)))
            return e;
        stamp(2);
        const bool halo = h->band.halo_rows > 0;
        if (R > 1 && (e = run_band_pull(h, n_pred, outbox_of_lower_neighbour, outbox_of_upper_neighbour,
                                        halo ? (const float*)edge_rows_of_lower_neighbour : nullptr,
                                        halo ? (const float*)edge_rows_of_upper_neighbour : nullptr)))
            return e;
        // from here on the particle count of the band is known on the device only: the host sizes the launches for an upper
        // bound, the kernels read the counts (h->band.dev_cnt)
        long long n_ub = R > 1 ? (long long)n_pred + h->band.est_recv : n_pred;
        if (n_ub > h->band.n_cap)
            n_ub = h->band.n_cap;
        h->band.dev_cnt = h->band.cnt;
        set_particle_counts(h, (int)n_ub, h->band.est_birth > 0 ? h->band.est_birth : 1);
        if (R > 1)
            h->hist0_valid = false; // the histogram of the first sort pass has to include the records that arrived
        h->pa_current = false;
        h->rec_valid = true;
        h->sorted_valid = false;
        h->meas_src = nullptr;
        if (measurement_band && measurement_band != h->meas)
        {
            if (on_device)
                h->meas_src = measurement_band; // the cell kernel copies it on the way
            else
                e = copy_in(h->meas, measurement_band, (size_t)h->C * sizeof(dogm_meas_cell), 0, h->stream);
        }
        h->band.halo_valid = (halo && R > 1) ? 1 : 0;
        if (!e && h->shift_grid_pending && h->shift.active && h->band.rows < h->band.G &&
            (h->shift.y_move > h->band.halo_rows || -h->shift.y_move > h->band.halo_rows))
            e = DOGM_ERR_INVALID_ARGUMENT;
        e = e ? e : run_assignment(h);
        e = e ? e : run_occupancy_update(h, dt);
        e = e ? e : run_persistent_weights(h, true);
        e = e ? e : run_born_scan(h);
        e = e ? e : run_band_publish_share(h, 0);
        h->band.dev_cnt = nullptr;
        if (e)
            return e; // (the other bands run into the bounded waits of their collectors)
        stamp(3);
    }
    if (stages & DOGM_BAND_STAGE_BIRTH)
    { // born mass of the whole grid; birth particles; joint CDF; this band's joint weight to all bands
        h->band.dev_cnt = h->band.cnt;
        e = run_band_collect_born(h);
        stamp(4);
        e = e ? e : run_birth_fill(h);
        e = e ? e : run_cdf(h);
        e = e ? e : run_band_publish_share(h, 1);
        h->band.dev_cnt = nullptr;
        if (e)
            return e;
        stamp(5);
    }
    if (stages & DOGM_BAND_STAGE_RESAMPLE)
    { // joint weight of the whole grid; this band's part of the draw
        h->band.dev_cnt = h->band.cnt;
        e = run_band_collect_weight(h);
        stamp(6);
        e = e ? e : run_resample_gather(h);
        h->band.dev_cnt = nullptr;
        if (e)
            return e;
        stamp(7);
        DOGM_CHECK(cudaMemcpyAsync(h->band.cnt_host, h->band.cnt, sizeof(BandCounts), cudaMemcpyDeviceToHost, h->stream));
        h->cycle++;
    }
    return 0;
}

extern "C" int dogm_band_set_profile(dogm_handle* h, int enable)
{
    if (!h || !h->band.enabled)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (enable && !h->band.stage_ev[0])
        for (int k = 0; k < 8; k++)
            DOGM_CHECK(cudaEventCreate(&h->band.stage_ev[k]));
    h->band.profile = enable != 0;
    return 0;
}

extern "C" int dogm_band_stage_times(const dogm_handle* h, float* out_ms7)
{
    if (!h || !h->band.enabled || !out_ms7)
        return DOGM_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < 7; k++)
        out_ms7[k] = h->band.stage_ms[k];
    return 0;
}

extern "C" int dogm_band_cycle_finish(dogm_handle* h, int* particles_out, int* sent_lo, int* sent_hi, double* born_total,
                                      double* weight_total)
{
    if (!h || !h->band.enabled || !h->band.cnt_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    const BandCounts c = *h->band.cnt_host;
    for (int k = 0; k < 7; k++)
    {
        h->band.stage_ms[k] = 0.0f;
        if (h->band.profile)
            cudaEventElapsedTime(&h->band.stage_ms[k], h->band.stage_ev[k], h->band.stage_ev[k + 1]);
    }
    set_particle_counts(h, c.n_out, c.B);
    if (h->band.est_forced <= 0)
    { // launch sizes of the next cycle: what this one needed plus a margin
        long long v = ((long long)c.n_lo + c.n_hi) * 5 / 4 + 16384;
        h->band.est_recv = (int)(v < 2ll * h->band.send_cap ? v : 2ll * h->band.send_cap);
        v = (long long)c.B * 9 / 8 + 4096;
        h->band.est_birth = (int)(v < h->band.b_cap ? v : h->band.b_cap);
        v = (long long)c.n_out * 17 / 16 + 8192;
        h->band.est_out = (int)(v < h->band.n_cap ? v : h->band.n_cap);
    }
    h->band.n_out = c.n_out;
    h->band.born_local = c.born_local;
    h->band.weight_local = c.weight_local;
    h->band.born_base = c.born_base;
    h->band.birth_slot_base = c.birth_slot_base;
    h->band.cdf_base = c.cdf_base;
    h->band.out_base = c.out_base;
    if (particles_out)
        *particles_out = c.n_out;
    if (sent_lo)
        *sent_lo = c.sent_lo;
    if (sent_hi)
        *sent_hi = c.sent_hi;
    if (born_total)
        *born_total = c.born_total;
    if (weight_total)
        *weight_total = c.weight_total;
    if (c.err & BAND_ERR_TIMEOUT)
        return DOGM_ERR_NOT_INITIALIZED; // a message of another band never arrived
    if (c.err)
        return DOGM_ERR_INVALID_ARGUMENT; // a capacity of the band configuration was exceeded
    return 0;
}

// Parity hook: loads a band with a given state instead of the first cycle's initialisation - the particles that lie in its rows
// (global coordinates; their cell indices are recomputed), the band's rows of the grid cells and the pose.
extern "C" int dogm_band_set_state(dogm_handle* h, int n, const float* state_xyvv, const float* weight, const unsigned char* associated,
                                   const dogm_grid_cell* grid_cells_band, float x, float y, float yaw)
{
    BAND_PROLOGUE();
    if (n < 0 || n > h->band.n_cap || (n > 0 && (!state_xyvv || !weight)))
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    set_particle_counts(h, n, 0);
    particle_set_assign(h->pa, h->pa.block, h->band.n_cap); // (the SoA block is laid out for the capacity)
    if (n > 0)
    {
        std::vector<int> idx((size_t)n);
        std::vector<unsigned char> as((size_t)n, 0);
        for (int i = 0; i < n; i++)
        {
            int px = (int)state_xyvv[4 * i], py = (int)state_xyvv[4 * i + 1];
            px = px < 0 ? 0 : (px > h->gs - 1 ? h->gs - 1 : px);
            py = py < 0 ? 0 : (py > h->band.G - 1 ? h->band.G - 1 : py);
            py -= h->band.row0;
            if (py < 0 || py >= h->band.rows)
                return DOGM_ERR_INVALID_ARGUMENT; // the particle does not lie in this band
            idx[i] = px + h->gs * py;
            if (associated)
                as[i] = associated[i];
        }
        DOGM_CHECK(cudaMemcpy(h->pa.state, state_xyvv, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
        DOGM_CHECK(cudaMemcpy(h->pa.idx, idx.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        DOGM_CHECK(cudaMemcpy(h->pa.weight, weight, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
        DOGM_CHECK(cudaMemcpy(h->pa.assoc, as.data(), (size_t)n, cudaMemcpyHostToDevice));
    }
    h->pa_current = true;
    h->rec_valid = false;
    h->sorted_valid = false;
    h->hist0_valid = false;
    if (grid_cells_band && (e = dogm_set_grid_cells(h, grid_cells_band, 0)))
        return e;
    h->position_x = x;
    h->position_y = y;
    h->yaw = yaw;
    h->first_pose_received = true;
    h->first_measurement_received = true;
    h->band.est_out = n + 8192 < h->band.n_cap ? n + 8192 : h->band.n_cap;
    return 0;
}

extern "C" int dogm_band_get_particles(dogm_handle* h, float* state_xyvv, int* cell_idx, float* weight, unsigned char* associated)
{
    BAND_PROLOGUE();
    e = ensure_soa(h);
    if (e)
        return e;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    const size_t n = (size_t)h->N;
    if (n == 0)
        return 0;
    if (state_xyvv)
        DOGM_CHECK(cudaMemcpy(state_xyvv, h->pa.state, n * sizeof(float4), cudaMemcpyDeviceToHost));
    if (cell_idx)
        DOGM_CHECK(cudaMemcpy(cell_idx, h->pa.idx, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (weight)
        DOGM_CHECK(cudaMemcpy(weight, h->pa.weight, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (associated)
        DOGM_CHECK(cudaMemcpy(associated, h->pa.assoc, n, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// instrumentation
// ---------------------------------------------------------------------------------------------------------
extern "C" void* dogm_get_stream(dogm_handle* h) { return h ? (void*)h->stream : nullptr; }
extern "C" uint64_t dogm_get_launch_count(const dogm_handle* h) { return h ? h->launch_count : 0; }

static void drain_timed(dogm_handle* h)
{
    cudaStreamSynchronize(h->stream);
    for (auto& t : h->timed)
    {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess)
        {
            h->acc_ms[t.id] += (double)ms;
            h->acc_launches[t.id] += 1;
        }
        h->event_pool.push_back(t.e0);
        h->event_pool.push_back(t.e1);
    }
    h->timed.clear();
}

extern "C" int dogm_debug_read(dogm_handle* h, const char* name, void* out_host, size_t bytes)
{
    if (!h || !name || !out_host)
        return DOGM_ERR_INVALID_ARGUMENT;
    const void* src = nullptr;
    size_t have = 0;
    if (!strcmp(name, "res_start"))
    {
        src = h->res_start;
        have = (size_t)div_up(h->N > 0 ? h->N : 1, kBlock) * sizeof(int);
    }
    else if (!strcmp(name, "bucket_start") && h->bucket.enabled)
    { // first position of every bucket in the grouping pass of the last assignment
        src = h->hist[0];
        have = (size_t)h->bucket.bins * sizeof(uint32_t);
    }
    else if (!strcmp(name, "bucket_org") && h->bucket.enabled)
    {
        src = h->bucket.org;
        have = (size_t)h->bucket.bins * sizeof(int);
    }
    else if (!strcmp(name, "bucket_samples") && h->bucket.enabled)
    {
        src = h->bucket.smp_raw;
        have = (size_t)h->bucket.n_spl * sizeof(int);
    }
    else if (!strcmp(name, "weight_total"))
    {
        src = &h->scal->weight_total;
        have = sizeof(double);
    }
    if (!strcmp(name, "phase_birth"))
    {
        DOGM_CHECK(cudaStreamSynchronize(h->stream));
        return debug_phase_read_cells(out_host, bytes);
    }
    if (!strcmp(name, "phase_chain") || !strcmp(name, "phase_resample"))
    {
        DOGM_CHECK(cudaStreamSynchronize(h->stream));
        return debug_phase_read(!strcmp(name, "phase_resample"), out_host, bytes);
    }
    if (!src || bytes > have)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    DOGM_CHECK(cudaMemcpy(out_host, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int dogm_trace_arm(dogm_handle* h, int enable)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaDeviceSynchronize());
    if (enable && !h->trace_buf)
        DOGM_CHECK(cudaMalloc(&h->trace_buf, kTraceSlots * sizeof(unsigned long long)));
    if (h->trace_buf)
        DOGM_CHECK(cudaMemset(h->trace_buf, 0xff, kTraceSlots * sizeof(unsigned long long)));
    unsigned long long* p = enable ? h->trace_buf : nullptr;
    int e = trace_bind_particles(p);
    e = e ? e : trace_bind_cells(p);
    e = e ? e : trace_bind_meas(p);
    return e;
}

extern "C" int dogm_trace_read(dogm_handle* h, uint64_t* out_start_ns, char* out_names, int capacity, int* out_count)
{
    if (!h || !out_start_ns || !out_names || !out_count || !h->trace_buf)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaStreamSynchronize(h->stream));
    unsigned long long host[kTraceSlots];
    DOGM_CHECK(cudaMemcpy(host, h->trace_buf, sizeof(host), cudaMemcpyDeviceToHost));
    DOGM_CHECK(cudaMemset(h->trace_buf, 0xff, sizeof(host)));
    int n = 0;
    for (int s = 0; s < kTraceSlots && n < capacity; s++)
    {
        if (host[s] == ~0ull)
            continue;
        out_start_ns[n] = host[s];
        snprintf(out_names + 48 * n, 48, "%s%s", kKernelNames[s / 2], (s & 1) ? "#2" : "");
        n++;
    }
    *out_count = n;
    return 0;
}

extern "C" int dogm_kernel_timing_enable(dogm_handle* h, int enable)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    drain_timed(h);
    h->timing = enable != 0;
    for (int k = 0; k < K_COUNT; k++)
    {
        h->acc_ms[k] = 0.0;
        h->acc_launches[k] = 0;
        h->acc_bytes[k] = 0.0;
    }
    return 0;
}

extern "C" int dogm_kernel_timing_read(dogm_handle* h, dogm_kernel_time* out, int capacity, int* out_count)
{
    if (!h || !out || !out_count)
        return DOGM_ERR_INVALID_ARGUMENT;
    drain_timed(h);
    int n = 0;
    for (int k = 0; k < K_COUNT && n < capacity; k++)
    {
        if (h->acc_launches[k] == 0)
            continue;
        memset(&out[n], 0, sizeof(dogm_kernel_time));
        strncpy(out[n].name, kKernelNames[k], sizeof(out[n].name) - 1);
        out[n].total_ms = h->acc_ms[k];
        out[n].launches = h->acc_launches[k];
        out[n].algorithmic_bytes = h->acc_bytes[k] / (double)h->acc_launches[k]; // average over the launches timed
        n++;
        h->acc_ms[k] = 0.0;
        h->acc_launches[k] = 0;
        h->acc_bytes[k] = 0.0;
    }
    *out_count = n;
    return 0;
}

static int ensure_timer(dogm_handle* h)
{
    if (!h->timer_ready)
    {
        DOGM_CHECK(cudaEventCreate(&h->timer_e0));
        DOGM_CHECK(cudaEventCreate(&h->timer_e1));
        h->timer_ready = true;
    }
    return 0;
}

extern "C" int dogm_timer_start(dogm_handle* h)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    int e = ensure_timer(h);
    if (e)
        return e;
    DOGM_CHECK(cudaEventRecord(h->timer_e0, h->stream));
    return 0;
}

extern "C" int dogm_timer_stop(dogm_handle* h)
{
    if (!h)
        return DOGM_ERR_INVALID_ARGUMENT;
    int e = ensure_timer(h);
    if (e)
        return e;
    DOGM_CHECK(cudaEventRecord(h->timer_e1, h->stream));
    return 0;
}

extern "C" int dogm_timer_elapsed_ms(dogm_handle* h, float* out_ms)
{
    if (!h || !out_ms || !h->timer_ready)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaEventSynchronize(h->timer_e1));
    DOGM_CHECK(cudaEventElapsedTime(out_ms, h->timer_e0, h->timer_e1));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// small runtime helpers for callers without a CUDA binding
// ---------------------------------------------------------------------------------------------------------
extern "C" int dogm_host_alloc_pinned(void** out, size_t bytes)
{
    if (!out)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaMallocHost(out, bytes ? bytes : 16));
    return 0;
}
extern "C" int dogm_host_free_pinned(void* p)
{
    DOGM_CHECK(cudaFreeHost(p));
    return 0;
}
extern "C" int dogm_device_alloc(void** out, size_t bytes)
{
    if (!out)
        return DOGM_ERR_INVALID_ARGUMENT;
    DOGM_CHECK(cudaMalloc(out, bytes ? bytes : 16));
    return 0;
}
extern "C" int dogm_device_free(void* p)
{
    DOGM_CHECK(cudaFree(p));
    return 0;
}
extern "C" int dogm_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes)
{
    DOGM_CHECK(cudaMemcpy(dst_device, src_host, bytes, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int dogm_memcpy_d2d(void* dst_device, const void* src_device, size_t bytes)
{ // unified addressing: also a peer copy when the two buffers live on different GPUs of this process
    DOGM_CHECK(cudaMemcpy(dst_device, src_device, bytes, cudaMemcpyDefault));
    return 0;
}
extern "C" int dogm_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes)
{
    DOGM_CHECK(cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int dogm_enable_peer_access(int device, int peer_device)
{ // lets `device` read and write `peer_device`'s memory directly (NVLink / NVSwitch); copies between the two then bypass the host
    if (device == peer_device)
        return 0;
    int can = 0;
    DOGM_CHECK(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can)
        return DOGM_ERR_UNSUPPORTED;
    int prev = 0;
    DOGM_CHECK(cudaGetDevice(&prev));
    DOGM_CHECK(cudaSetDevice(device));
    cudaError_t pe = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (pe == cudaErrorPeerAccessAlreadyEnabled)
    {
        cudaGetLastError();
        pe = cudaSuccess;
    }
    cudaSetDevice(prev);
    DOGM_CHECK(pe);
    return 0;
}
extern "C" int dogm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}
extern "C" int dogm_set_device(int device)
{
    DOGM_CHECK(cudaSetDevice(device));
    return 0;
}
extern "C" const char* dogm_b200_version(void)
{
    return "dogm_b200 0.1 (sm_100a, ABI " /* DOGM_B200_ABI_VERSION */ "1)";
}
