// Native in-process orchestrator of band mode (include/dogm_b200.h, "Band group"): one persistent host thread per band, each
// bound to its band's GPU, driving the dogm_band_* phases of a cycle; between the phases the threads meet at a spinning
// barrier and every thread adds up the bands' normaliser shares in band order (so all bands see bit-identical prefixes).
// The per-cycle cost of the orchestration itself is a handful of barrier crossings (microseconds), where a scripting-language
// thread pool needs tens of microseconds per phase and band.
//
// Two ways to run a cycle (dogm_band_group_set_mode):
//   device-paced (default): every band thread enqueues its band's whole cycle (dogm_band_cycle_enqueue) and waits once at the
//     end; what the bands tell each other - record counts, the records and halo rows themselves, the two normalisers - travels
//     GPU to GPU inside the streams (peer loads / stores, single-block kernels that wait for the messages), no host barrier;
//   host-paced: the phase calls, one synchronisation + one barrier per phase (the first cycle's initialisation always).
// The reference has nothing to compare with (single device, dogm.cu:39-43); the phases are those of DESIGN.md section 8.
#include "dogm_internal.cuh"

#include <atomic>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace
{

struct SpinBarrier
{ // sense-reversing barrier for a fixed number of threads; waits spin (a cycle has ~6 crossings, a few hundred microseconds apart)
    explicit SpinBarrier(int n_) : n(n_), waiting(0), sense(0) {}
    void wait()
    {
        const int my = sense.load(std::memory_order_acquire);
        if (waiting.fetch_add(1, std::memory_order_acq_rel) + 1 == n)
        {
            waiting.store(0, std::memory_order_relaxed);
            sense.store(my ^ 1, std::memory_order_release);
        }
        else
        {
            int spins = 0;
            while (sense.load(std::memory_order_acquire) == my)
                if (++spins > 4096)
                    std::this_thread::yield();
        }
    }
    const int n;
    std::atomic<int> waiting;
    std::atomic<int> sense;
};

} // namespace

struct dogm_band_group
{
    int n = 0;
    std::vector<dogm_handle*> bands;
    std::vector<int> devices;
    std::vector<std::thread> workers;
    SpinBarrier* barrier = nullptr;
    // command hand-over
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    uint64_t generation = 0;
    int finished = 0;
    bool quit = false;
    bool first = true;
    // the cycle's arguments and results
    const dogm_meas_cell* const* meas = nullptr;
    float x = 0, y = 0, yaw = 0, dt = 0;
    std::vector<int> sent_lo, sent_hi, counts;
    std::vector<double> share; // the normaliser share of every band in the running phase
    std::atomic<int> error{0};
    double born_total = 0.0, weight_total = 0.0;
    std::chrono::steady_clock::time_point stamp[6];
    std::vector<float> band_ms; // [band][5]: what each band itself spent in every phase (without the waiting at the barriers)
    int mode = DOGM_BAND_GROUP_HOST_PACED;
    bool linked = false;               // every band knows every band's mailbox and can reach it
    bool shared_device = false;        // at least two bands live on the same GPU
    std::vector<const void*> edge_lo, edge_hi; // the bands' edge rows of the previous free masses at the start of the cycle
};

namespace
{

void prefix_of(const std::vector<double>& v, int r, double* before, double* total)
{ // summed in band order: every band computes the same doubles
    double acc = 0.0, b = 0.0;
    for (size_t k = 0; k < v.size(); k++)
    {
        if ((int)k == r)
            b = acc;
        acc = acc + v[k];
    }
    *before = b;
    *total = acc;
}

void fail(dogm_band_group* g, int e)
{
    int expected = 0;
    if (e)
        g->error.compare_exchange_strong(expected, e);
}

void run_cycle(dogm_band_group* g, int r)
{
    dogm_handle* h = g->bands[r];
    const int R = g->n;
    const dogm_meas_cell* meas = g->meas ? g->meas[r] : nullptr;
    auto ok = [&]() { return g->error.load(std::memory_order_acquire) == 0; };
    auto mark = [&](int k) {
        if (r == 0)
            g->stamp[k] = std::chrono::steady_clock::now();
    };
    if (g->first)
    { // first cycle: the initial particles are distributed over the whole grid in proportion to the measured occupancy
        double m = 0.0;
        fail(g, dogm_band_init_masses(h, meas, 1, &m));
        g->share[r] = m;
        g->barrier->wait();
        double before, total;
        prefix_of(g->share, r, &before, &total);
        if (ok())
            fail(g, dogm_band_init_particles(h, before, total, nullptr));
        g->barrier->wait();
    }
    mark(0);
    auto t_begin = std::chrono::steady_clock::now();
    if (g->mode == DOGM_BAND_GROUP_DEVICE_PACED && g->linked)
    { // the whole cycle in one go; the bands meet inside their streams
        const bool halo = h->band.halo_rows > 0;
        const void* box_lo = r > 0 ? dogm_band_buffer(g->bands[r - 1], DOGM_BAND_SEND_HI) : nullptr;
        const void* box_hi = r + 1 < R ? dogm_band_buffer(g->bands[r + 1], DOGM_BAND_SEND_LO) : nullptr;
        const void* edge_lo = (halo && r > 0) ? g->edge_hi[r - 1] : nullptr;
        const void* edge_hi = (halo && r + 1 < R) ? g->edge_lo[r + 1] : nullptr;
        int e = 0;
        if (!g->shared_device)
            e = dogm_band_cycle_enqueue(h, DOGM_BAND_STAGE_ALL, meas, 1, g->x, g->y, g->yaw, g->dt, box_lo, box_hi, edge_lo, edge_hi);
        else
        { // bands that share a GPU: stage k of every band is enqueued before stage k + 1 of any (host barrier, no device sync)
            for (int stage = DOGM_BAND_STAGE_PREDICT; stage <= DOGM_BAND_STAGE_RESAMPLE; stage <<= 1)
            {
                if (e == 0)
                    e = dogm_band_cycle_enqueue(h, stage, meas, 1, g->x, g->y, g->yaw, g->dt, box_lo, box_hi, edge_lo, edge_hi);
                if (stage != DOGM_BAND_STAGE_RESAMPLE)
                    g->barrier->wait();
            }
        }
        int n_out = 0, a = 0, b = 0;
        double born = 0.0, weight = 0.0;
        const int e2 = dogm_band_cycle_finish(h, &n_out, &a, &b, &born, &weight); // (also drains the stream after a failed enqueue)
        fail(g, e ? e : e2);
        g->counts[r] = n_out;
        g->sent_lo[r] = a;
        g->sent_hi[r] = b;
        if (r == 0)
        {
            g->born_total = born;
            g->weight_total = weight;
        }
        { // own work per stage when the bands are profiled (dogm_band_set_profile; the waits for other bands are left out), zeros otherwise
            float st[7] = {0, 0, 0, 0, 0, 0, 0};
            dogm_band_stage_times(h, st);
            g->band_ms[(size_t)r * 5 + 0] = st[0];
            g->band_ms[(size_t)r * 5 + 1] = st[1] + st[3] + st[5]; // waiting for the other bands
            g->band_ms[(size_t)r * 5 + 2] = st[2];
            g->band_ms[(size_t)r * 5 + 3] = st[4];
            g->band_ms[(size_t)r * 5 + 4] = st[6];
        }
        for (int k = 1; k <= 5; k++)
            mark(k);
        return;
    }
    auto lap = [&](int k) {
        const auto now = std::chrono::steady_clock::now();
        g->band_ms[(size_t)r * 5 + k] = std::chrono::duration<float, std::milli>(now - t_begin).count();
    };
    auto restart = [&]() { t_begin = std::chrono::steady_clock::now(); };
    int a = 0, b = 0;
    if (ok())
        fail(g, dogm_band_predict(h, g->x, g->y, g->yaw, g->dt, &a, &b));
    g->sent_lo[r] = a;
    g->sent_hi[r] = b;
    lap(0);
    g->barrier->wait();
    restart();
    mark(1);
    if (ok())
    { // what the neighbours sent towards this band, and their edge rows of the previous free masses
        const int n_lo = r > 0 ? g->sent_hi[r - 1] : 0;
        const int n_hi = r + 1 < R ? g->sent_lo[r + 1] : 0;
        const bool halo = h->band.halo_rows > 0;
        const void* box_lo = n_lo ? dogm_band_buffer(g->bands[r - 1], DOGM_BAND_SEND_HI) : nullptr;
        const void* box_hi = n_hi ? dogm_band_buffer(g->bands[r + 1], DOGM_BAND_SEND_LO) : nullptr;
        const void* edge_lo = (halo && r > 0) ? dogm_band_buffer(g->bands[r - 1], DOGM_BAND_EDGE_HI) : nullptr;
        const void* edge_hi = (halo && r + 1 < R) ? dogm_band_buffer(g->bands[r + 1], DOGM_BAND_EDGE_LO) : nullptr;
        fail(g, dogm_band_receive(h, box_lo, n_lo, box_hi, n_hi, edge_lo, edge_hi));
    }
    lap(1);
    g->barrier->wait(); // every band has its halo rows before any band's cell kernel swaps the free masses
    mark(2);
    restart();
    double v = 0.0;
    if (ok())
        fail(g, dogm_band_update(h, meas, 1, g->dt, h->band.halo_rows > 0 ? 1 : 0, &v));
    g->share[r] = v;
    lap(2);
    g->barrier->wait();
    mark(3);
    double before, total;
    prefix_of(g->share, r, &before, &total);
    if (r == 0)
        g->born_total = total;
    g->barrier->wait(); // (share[] is rewritten next)
    restart();
    v = 0.0;
    if (ok())
        fail(g, dogm_band_birth(h, before, total, &v));
    g->share[r] = v;
    lap(3);
    g->barrier->wait();
    mark(4);
    restart();
    prefix_of(g->share, r, &before, &total);
    if (r == 0)
        g->weight_total = total;
    int n_out = 0;
    if (ok())
        fail(g, dogm_band_resample(h, before, total, &n_out));
    g->counts[r] = n_out;
    lap(4);
    g->barrier->wait();
    mark(5);
}

void worker_main(dogm_band_group* g, int r)
{
    cudaSetDevice(g->devices[r]);
    uint64_t seen = 0;
    for (;;)
    {
        {
            std::unique_lock<std::mutex> lk(g->m);
            g->cv_go.wait(lk, [&] { return g->quit || g->generation != seen; });
            if (g->quit)
                return;
            seen = g->generation;
        }
        run_cycle(g, r);
        {
            std::lock_guard<std::mutex> lk(g->m);
            if (++g->finished == g->n)
                g->cv_done.notify_all();
        }
    }
}

} // namespace

extern "C" int dogm_band_group_create(dogm_handle* const* bands, int n_bands, dogm_band_group** out)
{
    if (!bands || n_bands < 1 || !out)
        return DOGM_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < n_bands; r++)
        if (!bands[r] || !bands[r]->band.enabled)
            return DOGM_ERR_INVALID_ARGUMENT;
    dogm_band_group* g = new dogm_band_group();
    g->n = n_bands;
    g->bands.assign(bands, bands + n_bands);
    for (int r = 0; r < n_bands; r++)
        g->devices.push_back(bands[r]->device);
    // neighbouring bands on different GPUs talk over NVLink: without peer access every copy would be staged through the host
    for (int r = 0; r + 1 < n_bands; r++)
        if (g->devices[r] != g->devices[r + 1])
        {
            dogm_enable_peer_access(g->devices[r], g->devices[r + 1]);
            dogm_enable_peer_access(g->devices[r + 1], g->devices[r]);
        }
    g->sent_lo.assign(n_bands, 0);
    g->sent_hi.assign(n_bands, 0);
    g->counts.assign(n_bands, 0);
    g->share.assign(n_bands, 0.0);
    g->band_ms.assign((size_t)n_bands * 5, 0.0f);
    g->edge_lo.assign(n_bands, nullptr);
    g->edge_hi.assign(n_bands, nullptr);
    // device-paced cycles: every band stores its normaliser shares into every band's mailbox, so every GPU has to reach all
    // the others; if it cannot (or there are more bands than a mailbox has entries), the group stays host-paced
    if (n_bands <= dogm_b200::kMaxBands)
    {
        bool reach = true;
        for (int a = 0; a < n_bands && reach; a++)
            for (int b = 0; b < n_bands && reach; b++)
                if (g->devices[a] != g->devices[b] && dogm_enable_peer_access(g->devices[a], g->devices[b]) != 0)
                    reach = false;
        if (reach)
        {
            std::vector<void*> mails(n_bands);
            for (int r = 0; r < n_bands; r++)
                mails[r] = dogm_band_mailbox(bands[r]);
            int prev = 0;
            cudaGetDevice(&prev);
            bool ok = true;
            for (int r = 0; r < n_bands && ok; r++)
            {
                cudaSetDevice(g->devices[r]);
                ok = dogm_band_link(bands[r], r, n_bands, mails.data()) == 0;
            }
            cudaSetDevice(prev);
            g->linked = ok;
        }
    }
    for (int a = 0; a < n_bands; a++)
        for (int b = a + 1; b < n_bands; b++)
            if (g->devices[a] == g->devices[b])
                g->shared_device = true;
    {
        const char* env = getenv("DOGM_BAND_HOST_PACED");
        g->mode = (g->linked && !(env && env[0] == '1')) ? DOGM_BAND_GROUP_DEVICE_PACED : DOGM_BAND_GROUP_HOST_PACED;
    }
    g->barrier = new SpinBarrier(n_bands);
    for (int r = 0; r < n_bands; r++)
        g->workers.emplace_back(worker_main, g, r);
    *out = g;
    return 0;
}

extern "C" void dogm_band_group_destroy(dogm_band_group* g)
{
    if (!g)
        return;
    {
        std::lock_guard<std::mutex> lk(g->m);
        g->quit = true;
    }
    g->cv_go.notify_all();
    for (std::thread& t : g->workers)
        t.join();
    delete g->barrier;
    delete g;
}

extern "C" int dogm_band_group_update(dogm_band_group* g, const dogm_meas_cell* const* measurement_bands, float new_x, float new_y,
                                      float new_yaw, float dt, int* particles_out, dogm_band_cycle_info* info)
{
    if (!g)
        return DOGM_ERR_INVALID_ARGUMENT;
    if (g->first && !measurement_bands)
        return DOGM_ERR_INVALID_ARGUMENT;
    g->meas = measurement_bands;
    g->x = new_x;
    g->y = new_y;
    g->yaw = new_yaw;
    g->dt = dt;
    g->error.store(0);
    for (int r = 0; r < g->n; r++)
    { // (the cell kernel of a cycle swaps a band's free-mass buffers: the neighbours must see this cycle's "previous" rows)
        g->edge_lo[r] = dogm_band_buffer(g->bands[r], DOGM_BAND_EDGE_LO);
        g->edge_hi[r] = dogm_band_buffer(g->bands[r], DOGM_BAND_EDGE_HI);
    }
    {
        std::unique_lock<std::mutex> lk(g->m);
        g->finished = 0;
        g->generation++;
        g->cv_go.notify_all();
        g->cv_done.wait(lk, [&] { return g->finished == g->n; });
    }
    const int e = g->error.load();
    if (e == 0)
        g->first = false;
    if (particles_out)
        for (int r = 0; r < g->n; r++)
            particles_out[r] = g->counts[r];
    if (info)
    {
        info->born_total = g->born_total;
        info->weight_total = g->weight_total;
        long long sent = 0;
        for (int r = 0; r < g->n; r++)
            sent += g->sent_lo[r] + g->sent_hi[r];
        info->migrated = sent;
        for (int k = 0; k < 5; k++)
            info->phase_ms[k] = std::chrono::duration<float, std::milli>(g->stamp[k + 1] - g->stamp[k]).count();
    }
    return e;
}

extern "C" int dogm_band_group_mark_initialized(dogm_band_group* g)
{ // the bands were loaded with dogm_band_set_state: no first-cycle initialisation
    if (!g)
        return DOGM_ERR_INVALID_ARGUMENT;
    g->first = false;
    return 0;
}

extern "C" int dogm_band_group_set_mode(dogm_band_group* g, int mode)
{
    if (!g || (mode != DOGM_BAND_GROUP_HOST_PACED && mode != DOGM_BAND_GROUP_DEVICE_PACED))
        return DOGM_ERR_INVALID_ARGUMENT;
    if (mode == DOGM_BAND_GROUP_DEVICE_PACED && !g->linked)
        return DOGM_ERR_UNSUPPORTED; // some band's GPU cannot reach another band's memory
    g->mode = mode;
    return 0;
}

extern "C" int dogm_band_group_get_mode(const dogm_band_group* g)
{
    return g ? g->mode : DOGM_ERR_INVALID_ARGUMENT;
}

extern "C" int dogm_band_group_band_times(const dogm_band_group* g, float* out_ms)
{ // [band][5]: predict, exchange + append, update, birth + CDF, resample - the band's own time in the last cycle
    if (!g || !out_ms)
        return DOGM_ERR_INVALID_ARGUMENT;
    for (size_t k = 0; k < g->band_ms.size(); k++)
        out_ms[k] = g->band_ms[k];
    return 0;
}
