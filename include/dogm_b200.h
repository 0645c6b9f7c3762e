/*
 * dogm_b200.h — C ABI of the B200-native dynamic occupancy grid map (DOGM) update cycle.
 *
 * This is the drop-in boundary for the reference's per-cycle path.  The reference has no FFI layer:
 * callers link its static library and use the C++ class dogm::DOGM (reference dogm/include/dogm/dogm.h:20-199).
 * Every entry point below replaces one member of that class (file:line cited per entry) and uses plain
 * pointers and sizes only.  The header-only C++ class in include/dogm/dogm.h re-creates dogm::DOGM on top
 * of this ABI, so a reference user switches by changing the include path and linking libdogm_b200.so.
 *
 * Conventions
 *   - every function returning int returns 0 on success or a cudaError_t value (> 0) / DOGM_ERR_* (< 0);
 *     like the reference's CHECK_ERROR (cuda_utils.h:13-24) the library also prints
 *     "GPU Kernel Error: <msg> <file> <line>" and carries on where the reference carries on.
 *   - not thread-safe; one handle binds to the CUDA device that is current at dogm_create (dogm.cu:39-43).
 *   - layouts of dogm_params / dogm_grid_cell / dogm_meas_cell and of the particle block are bit-identical to
 *     DOGM::Params (dogm.h:26-60), GridCell / MeasurementCell / ParticlesSoA (dogm_types.h:13-144).
 *   - there is no CPU fallback: without a CUDA device dogm_create fails with a cudaError_t.
 */
#ifndef DOGM_B200_H
#define DOGM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOGM_B200_ABI_VERSION 1

#define DOGM_ERR_INVALID_ARGUMENT (-1)
#define DOGM_ERR_NOT_INITIALIZED (-2)
#define DOGM_ERR_UNSUPPORTED (-3)

/* == DOGM::Params, dogm.h:26-60 (11 fields, same order, same types) */
typedef struct dogm_params
{
    float size;                          /* grid edge length [m] */
    float resolution;                    /* cell edge length [m/cell] */
    int particle_count;                  /* persistent particles N */
    int new_born_particle_count;         /* birth particles B */
    float persistence_prob;              /* p_S */
    float stddev_process_noise_position; /* [cells] */
    float stddev_process_noise_velocity; /* [cells/s] */
    float birth_prob;                    /* p_B */
    float stddev_velocity;               /* birth velocity sigma */
    float init_max_velocity;             /* first-cycle velocity range */
    float freespace_discount;            /* alpha; per-cycle factor is alpha^dt */
} dogm_params;

/* == GridCell, dogm_types.h:13-33 (64 bytes, row-major idx = x + grid_size * y) */
typedef struct dogm_grid_cell
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
} dogm_grid_cell;

/* == MeasurementCell, dogm_types.h:35-41 (16 bytes) */
typedef struct dogm_meas_cell
{
    float free_mass;
    float occ_mass;
    float likelihood;
    float p_A;
} dogm_meas_cell;

/* Particle block == ParticlesSoA::memory_block, dogm_types.h:69-85,136-142:
 *   one block of n * 28 bytes: float4 state[n] (x, y, vx, vy) | int grid_cell_idx[n] | float weight[n] | bool associated[n]
 *   (the tail n*3 bytes of the block are padding, as in the reference where sizeof(Particle) == 28). */
#define DOGM_PARTICLE_BLOCK_BYTES(n) ((size_t)(n) * 28u)
#define DOGM_PARTICLE_STATE_OFFSET(n) ((size_t)0)
#define DOGM_PARTICLE_IDX_OFFSET(n) ((size_t)(n) * 16u)
#define DOGM_PARTICLE_WEIGHT_OFFSET(n) ((size_t)(n) * 20u)
#define DOGM_PARTICLE_ASSOC_OFFSET(n) ((size_t)(n) * 24u)

typedef struct dogm_handle dogm_handle;

/* ------------------------------------------------------------------------------------------------------------
 * Life cycle
 * ---------------------------------------------------------------------------------------------------------- */

/* DOGM::DOGM(const Params&) + DOGM::initialize(), dogm.cu:33-71,95-113.
 * Allocates every buffer once (no allocation happens in any later call), zero-fills the particle sets
 * (the reference leaves them uninitialised), cells: masses 0 / indices -1, measurement cells (0,0,1,1). */
int dogm_create(const dogm_params* params, dogm_handle** out);

/* DOGM::~DOGM(), dogm.cu:73-93 */
void dogm_destroy(dogm_handle* h);

/* ------------------------------------------------------------------------------------------------------------
 * The cycle
 * ---------------------------------------------------------------------------------------------------------- */

/* DOGM::updateGrid(measurement_grid, new_x, new_y, new_yaw, dt, device), dogm.cu:115-131.
 * measurement_grid may be NULL (the reference tolerates the failing copy, dogm_spec.cpp:32; here it means
 * "keep the previous measurement grid").  Returns after the whole cycle has finished on the device
 * (the reference ends with cudaDeviceSynchronize, dogm.cu:130). */
int dogm_update_grid(dogm_handle* h, const dogm_meas_cell* measurement_grid, float new_x, float new_y,
                     float new_yaw, float dt, int on_device);

/* Same cycle without the final synchronisation (stream-ordered; the host copy of a host measurement grid is
 * taken before returning).  Pair with dogm_synchronize. */
int dogm_update_grid_async(dogm_handle* h, const dogm_meas_cell* measurement_grid, float new_x, float new_y,
                           float new_yaw, float dt, int on_device);
int dogm_synchronize(dogm_handle* h);

/* The eight public stage methods, dogm.h:148-156 / dogm.cu:217-423.  They run on the handle's stream and
 * synchronise before returning.  Where this implementation fuses work across the reference's stage borders
 * the observable state after the *last* stage of a cycle is the reference's; DESIGN.md lists what each stage
 * leaves behind. */
int dogm_update_measurement_grid(dogm_handle* h, const dogm_meas_cell* measurement_grid, int on_device);
                                                               /* dogm.cu:205-215 (private in the reference) */
int dogm_update_pose(dogm_handle* h, float new_x, float new_y, float new_yaw);
                                                               /* dogm.cu:161-203 (private in the reference): records the
                                                                  pose and arms the ego-motion shift that the next
                                                                  prediction / occupancy-update stages apply */
int dogm_initialize_particles(dogm_handle* h);                 /* dogm.cu:217-241 */
int dogm_particle_prediction(dogm_handle* h, float dt);        /* dogm.cu:243-260 */
int dogm_particle_assignment(dogm_handle* h);                  /* dogm.cu:262-281 */
int dogm_grid_cell_occupancy_update(dogm_handle* h, float dt); /* dogm.cu:283-298 */
int dogm_update_persistent_particles(dogm_handle* h);          /* dogm.cu:300-321 */
int dogm_initialize_new_particles(dogm_handle* h);             /* dogm.cu:323-348 */
int dogm_statistical_moments(dogm_handle* h);                  /* dogm.cu:350-384 */
int dogm_resampling(dogm_handle* h);                           /* dogm.cu:386-423, plus the publish of dogm.cu:128 */

/* ------------------------------------------------------------------------------------------------------------
 * Band mode: one large grid shared by several handles (one per GPU), each owning a band of rows
 *
 * The reference runs on one device (dogm.cu:39-43).  Here a G x G grid can be split into bands of whole rows
 * (row-major idx = x + G * y makes a band a contiguous range of cells): a band handle holds the cells of its rows and
 * the particles that currently lie there.  One cycle is driven in phases by an orchestrator, which moves three things
 * between the bands (device-to-device copies inside a process, NVLink peer copies or NCCL between processes):
 *   1. the particles that crossed a band edge during prediction (32-byte records, dogm_band_buffer SEND -> RECV),
 *   2. per band one double per normaliser: born mass (birth slots are handed out over the whole grid) and joint
 *      weight (resampling draws over the whole grid); every band gets the sum of the bands in front of it and the total,
 *   3. |y_move| rows of the previous free masses next to a band edge when the ego-motion shift moves the grid in y.
 * Everything else (sort, per-cell sums, occupancy update, weights, moments, CDF, gather) is local to a band.
 * Particle coordinates stay global; particle counts per band vary from cycle to cycle within the capacities.
 * With a single band covering the whole grid the phases reproduce dogm_update_grid bit for bit.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct dogm_band_config
{
    int row0;               /* first row of the band */
    int rows;               /* number of rows */
    int particle_capacity;  /* persistent particles the band can hold (its share plus head-room for drift) */
    int birth_capacity;     /* birth particles the band can hold */
    int exchange_capacity;  /* particles per edge and cycle that may leave / arrive */
    int halo_rows;          /* largest |y_move| of an ego-motion shift */
    uint64_t seed_salt;     /* added to the seed for per-particle noise: bands must not share their noise streams */
} dogm_band_config;

#define DOGM_BAND_SEND_LO 0 /* records leaving through the lower edge (towards smaller rows) */
#define DOGM_BAND_SEND_HI 1
#define DOGM_BAND_RECV_LO 2 /* records arriving through the lower edge */
#define DOGM_BAND_RECV_HI 3
#define DOGM_BAND_HALO_LO 4 /* halo_rows rows below the band: filled from the lower neighbour's DOGM_BAND_EDGE_HI */
#define DOGM_BAND_HALO_HI 5
#define DOGM_BAND_EDGE_LO 6 /* the band's own first halo_rows rows of the previous free masses */
#define DOGM_BAND_EDGE_HI 7 /* its last halo_rows rows */

/* params describe the WHOLE grid (size, resolution, particle counts of all bands together) */
int dogm_create_band(const dogm_params* params, const dogm_band_config* band, dogm_handle** out);
/* device address of an exchange buffer (see the DOGM_BAND_* names); sizes: exchange_capacity * 32 bytes for the particle
 * boxes, halo_rows * G * 4 bytes for the row buffers */
void* dogm_band_buffer(dogm_handle* h, int which);
/* current persistent / birth particle counts of the band */
int dogm_band_counts(dogm_handle* h, int* particles, int* birth_particles);
/* first cycle, step 1: takes the band's measurement grid (rows * G cells), returns the band's initial mass
 * (initializeParticles, dogm.cu:217-241: copyMassesKernel + scan) */
int dogm_band_init_masses(dogm_handle* h, const dogm_meas_cell* measurement_band, int on_device, double* mass_local);
/* first cycle, step 2: the band creates its share of the particles (initParticlesKernel1/2); mass_before = sum of
 * the masses of the bands in front of this one */
int dogm_band_init_particles(dogm_handle* h, double mass_before, double mass_total, int* particles);
/* every cycle, step 1: ego-motion bookkeeping (updatePose) + prediction; fills the SEND boxes (in the order of the senders' slots:
 * runs are bit-reproducible for any number of bands) and returns their counts */
int dogm_band_predict(dogm_handle* h, float new_x, float new_y, float new_yaw, float dt, int* send_lo, int* send_hi);
/* step 2: the records the orchestrator put into the RECV boxes join the band's particles */
int dogm_band_append(dogm_handle* h, int recv_lo, int recv_hi);
/* The whole exchange step in one call: copies `recv_lo` records from the lower neighbour's DOGM_BAND_SEND_HI box and `recv_hi`
 * from the upper neighbour's DOGM_BAND_SEND_LO box into this band's inboxes, the neighbours' edge rows (DOGM_BAND_EDGE_HI of the
 * lower, DOGM_BAND_EDGE_LO of the upper neighbour; NULL = none) into its halo rows, and appends the records - all on the band's
 * stream, one synchronisation.  The source addresses may lie on another GPU of this process (see dogm_enable_peer_access). */
int dogm_band_receive(dogm_handle* h, const void* outbox_of_lower_neighbour, int recv_lo, const void* outbox_of_upper_neighbour,
                      int recv_hi, const void* edge_rows_of_lower_neighbour, const void* edge_rows_of_upper_neighbour);
/* step 3: assignment, occupancy update (with the halo rows when halo_valid), persistent weights; returns the band's born
 * mass.  measurement_band may be NULL after the first cycle's dogm_band_init_masses (keeps that grid). */
int dogm_band_update(dogm_handle* h, const dogm_meas_cell* measurement_band, int on_device, float dt, int halo_valid,
                     double* born_local);
/* step 4: birth particles of the band's cells (their slots are numbered over the whole grid) and the band's joint weight
 * CDF; returns the band's joint weight */
int dogm_band_birth(dogm_handle* h, double born_before, double born_total, double* weight_local);
/* step 5: the band draws the output slots whose offsets fall into its part of the global CDF; returns how many */
int dogm_band_resample(dogm_handle* h, double weight_before, double weight_total, int* particles_out);
/* the host arithmetic of steps 4 and 5 on its own (no device needed): which birth slots / output slots of the whole grid a
 * band gets, given the mass / weight in front of it, its own share and the total.  Consecutive bands tile [0, total) when
 * every band's `before` is the running sum before[r + 1] = before[r] + local[r]. */
int dogm_band_slot_range(double mass_before, double mass_local, double mass_total, int slots_total, int* first, int* past);
int dogm_band_output_range(uint64_t seed, uint32_t cycle, int resample_mode, long long n_glob, double weight_before,
                           double weight_local, double weight_total, long long* first, long long* past);
/* read-out of the band's current particles (n = dogm_band_counts): any pointer may be NULL */
/* Band group: the native in-process orchestrator of the phases above for the GPUs of one box.  One persistent host thread per
 * band (bound to the band's GPU) runs the phases, the threads meet at a barrier where the bands depend on each other, the
 * normaliser shares are added up in band order, records and halo rows move by direct GPU-to-GPU copies (peer access is enabled
 * between neighbouring bands' GPUs).  The group borrows the handles (destroy the group first). */
typedef struct dogm_band_group dogm_band_group;
typedef struct dogm_band_cycle_info
{
    double born_total;   /* born mass of the whole grid in this cycle */
    double weight_total; /* joint weight of the whole grid before resampling */
    long long migrated;  /* particles that changed band */
    float phase_ms[5];   /* host wall clock: predict, exchange + append, update, birth + CDF, resample */
} dogm_band_cycle_info;
int dogm_band_group_create(dogm_handle* const* bands, int n_bands, dogm_band_group** out);
void dogm_band_group_destroy(dogm_band_group* g);
/* One cycle of the whole grid.  measurement_bands[r]: the rows of the measurement grid band r owns, on band r's GPU (required in
 * the first cycle; NULL entries later keep the previous grid).  particles_out[r] (may be NULL): band r's particle count after
 * resampling.  Blocking: returns when every band has finished the cycle. */
int dogm_band_group_update(dogm_band_group* g, const dogm_meas_cell* const* measurement_bands, float new_x, float new_y,
                           float new_yaw, float dt, int* particles_out, dogm_band_cycle_info* info);

/* How a group drives its cycles.  DEVICE_PACED (the default when every band's GPU can reach every other's memory): each band
 * enqueues its whole cycle at once; the particles that change band, the halo rows and the two normalisers travel GPU to GPU -
 * a band's kernels read its neighbours' send boxes directly, every band stores its normaliser share into every band's mailbox
 * and single-block kernels wait for the shares in the stream - and the host synchronises once, at the end of the cycle.
 * HOST_PACED: the phase calls above, one host synchronisation and one barrier per phase.  Same results bit for bit. */
#define DOGM_BAND_GROUP_HOST_PACED 0
#define DOGM_BAND_GROUP_DEVICE_PACED 1
int dogm_band_group_set_mode(dogm_band_group* g, int mode);
int dogm_band_group_get_mode(const dogm_band_group* g);

/* The building blocks of the device-paced cycle, for an orchestrator of its own (one call per band and cycle each):
 * dogm_band_link tells a band its rank and all bands' mailboxes (dogm_band_mailbox; every band's GPU needs peer access to the
 * others'); dogm_band_cycle_enqueue enqueues the cycle (the neighbours' DOGM_BAND_SEND_* boxes and DOGM_BAND_EDGE_* rows are read
 * in place by this band's kernels; the first cycle still runs through dogm_band_init_*); dogm_band_cycle_finish waits for it
 * and returns the band's new particle count (DOGM_ERR_INVALID_ARGUMENT: a capacity was exceeded).
 * `stages`: DOGM_BAND_STAGE_ALL enqueues the whole cycle.  Each of the four stages ends by publishing a message to the other
 * bands and the next begins by waiting for theirs inside the stream.  Bands that share ONE GPU must be enqueued stage by
 * stage - stage k of every band before stage k + 1 of any (their streams may share a hardware queue, and a waiting kernel
 * must never sit in front of the publisher it waits for); bands on GPUs of their own take DOGM_BAND_STAGE_ALL. */
#define DOGM_BAND_STAGE_PREDICT 1  /* prediction, outbox, record counts -> neighbours */
#define DOGM_BAND_STAGE_UPDATE 2   /* neighbours' records and halo rows, sort, per-cell sums, occupancy update, born mass -> all */
#define DOGM_BAND_STAGE_BIRTH 4    /* birth particles, joint CDF, joint weight -> all */
#define DOGM_BAND_STAGE_RESAMPLE 8 /* this band's part of the global draw */
#define DOGM_BAND_STAGE_ALL 15
void* dogm_band_mailbox(dogm_handle* h);
int dogm_band_link(dogm_handle* h, int rank, int n_bands, void* const* mailboxes);
int dogm_band_cycle_enqueue(dogm_handle* h, int stages, const dogm_meas_cell* measurement_band, int on_device, float new_x,
                            float new_y, float new_yaw, float dt, const void* outbox_of_lower_neighbour,
                            const void* outbox_of_upper_neighbour, const void* edge_rows_of_lower_neighbour,
                            const void* edge_rows_of_upper_neighbour);
int dogm_band_cycle_finish(dogm_handle* h, int* particles_out, int* sent_lo, int* sent_hi, double* born_total,
                           double* weight_total);
/* Profiling of the device-paced cycle: with enable != 0 events are recorded on the band's stream around the three waits for the
 * other bands' messages (they interrupt the launch chain, so a profiled cycle is a little slower); dogm_band_stage_times returns
 * seven device times of the last profiled cycle, in stream order: [0] prediction + outbox + record counts published,
 * [1] WAIT for the neighbours' counts, [2] pull + sort + per-cell sums + occupancy update + born mass published, [3] WAIT for the
 * born masses, [4] birth particles + CDF + joint weight published, [5] WAIT for the joint weights, [6] resampling.  The even
 * entries are the band's own work (what the band edges have to balance), the odd ones what it lost waiting for slower bands. */
int dogm_band_set_profile(dogm_handle* h, int enable);
int dogm_band_stage_times(const dogm_handle* h, float* out_ms7);

/* out_ms[band * 5 + phase]: what every band itself spent in the phases of the last cycle (host wall clock, without the waiting
 * at the barriers) - the input for placing the band edges */
int dogm_band_group_band_times(const dogm_band_group* g, float* out_ms);

int dogm_band_get_particles(dogm_handle* h, float* state_xyvv, int* cell_idx, float* weight, unsigned char* associated);
/* Parity hooks: a band loaded with a given state (the particles that lie in its rows, in global coordinates; its rows of the grid
 * cells; the pose) instead of the first cycle's initialisation, and a group told that its bands are initialised. */
int dogm_band_set_state(dogm_handle* h, int n, const float* state_xyvv, const float* weight, const unsigned char* associated,
                        const dogm_grid_cell* grid_cells_band, float x, float y, float yaw);
int dogm_band_group_mark_initialized(dogm_band_group* g);

/* ------------------------------------------------------------------------------------------------------------
 * Read-out (blocking device-to-host copies, dogm.cu:133-159, dogm.h:111-139)
 * ---------------------------------------------------------------------------------------------------------- */
int dogm_get_grid_cells(dogm_handle* h, dogm_grid_cell* out_host);        /* getGridCells: grid_cell_count * 64 B */
/* getGridCells without stopping the filter (the reference's demo loop reads all cells after every updateGrid,
 * demo/main.cpp:94-110 -> dogm.cu:133-141: 64 B per cell over PCIe, ten times the cycle itself at the 1200^2 size).
 * _begin enqueues the copy of the cells of the cycle enqueued last into out_host (page-locked memory, or the copy is not
 * asynchronous) on a copy stream of the handle and returns; the copy then runs under the dogm_update_grid* calls that follow:
 * from the first _begin on the handle keeps two GridCell buffers and the cell kernel alternates between them
 * (dogm_get_device_ptrs / the C++ class's grid_cell_array follow).  _wait returns when every copy begun so far has arrived.
 *   update(n); begin(n, buf[n & 1]); ... update(n + 1); wait(); use buf[n & 1]; begin(n + 1, buf[(n + 1) & 1]); ...
 * Not available on band handles (DOGM_ERR_UNSUPPORTED). */
int dogm_get_grid_cells_begin(dogm_handle* h, dogm_grid_cell* out_host);
int dogm_get_grid_cells_wait(dogm_handle* h);
int dogm_get_measurement_cells(dogm_handle* h, dogm_meas_cell* out_host); /* getMeasurementCells: count * 16 B */
int dogm_get_particles(dogm_handle* h, void* out_block);                  /* getParticles: particle block of N */
int dogm_get_grid_size(const dogm_handle* h);                             /* getGridSize */
int dogm_get_grid_cell_count(const dogm_handle* h);
int dogm_get_particle_count(const dogm_handle* h);
int dogm_get_new_born_particle_count(const dogm_handle* h);
float dogm_get_resolution(const dogm_handle* h);
float dogm_get_position_x(const dogm_handle* h);
float dogm_get_position_y(const dogm_handle* h);
float dogm_get_yaw(const dogm_handle* h);

/* The reference's public device buffers (dogm.h:159-191).  particle_array / particle_array_next swap roles
 * inside a cycle (pointer swap instead of the reference's 28*N-byte copy, dogm.cu:128), so re-read this
 * struct after every call that runs a stage. */
typedef struct dogm_device_ptrs
{
    dogm_grid_cell* grid_cell_array;
    void* particle_array;           /* particle block of N */
    void* particle_array_next;      /* particle block of N */
    void* birth_particle_array;     /* particle block of B */
    dogm_meas_cell* meas_cell_array;
    float* weight_array;            /* N */
    float* born_masses_array;       /* grid_cell_count */
    int* resampled_idx_array;       /* N ancestors of the last resampling (a temporary in the reference, dogm.cu:413) */
    double* joint_weight_accum;     /* N + B running sum of [weight_array, birth weights] of the last resampling */
    int* cell_start_array;          /* grid_cell_count, -1 for empty cells (== GridCell.start_idx) */
    int* cell_end_array;            /* grid_cell_count, valid where cell_start_array >= 0 */
} dogm_device_ptrs;
int dogm_get_device_ptrs(dogm_handle* h, dogm_device_ptrs* out);

/* ------------------------------------------------------------------------------------------------------------
 * Options and parity hooks (no equivalent in the reference: it hard-codes XORWOW seed 123456, dogm.cu:101)
 * ---------------------------------------------------------------------------------------------------------- */

/* How the N sorted resampling offsets are produced */
#define DOGM_RESAMPLE_SYSTEMATIC 0 /* (i + u0) * total / N with one Philox draw per cycle (default) */
#define DOGM_RESAMPLE_STRATIFIED 1 /* (i + u_i) * total / N with one Philox draw per particle */
#define DOGM_RESAMPLE_INJECTED 2   /* total * a_i with a caller-supplied ascending a_i in [0,1): the reference's
                                      sorted multinomial draws (resampling.cu:17-47) fed from outside */

/* Source of the Gaussian / uniform noise of prediction, birth and first-cycle initialisation */
#define DOGM_NOISE_PHILOX 0   /* counter-based Philox4x32-10, counter = (slot, stage, cycle) */
#define DOGM_NOISE_INJECTED 1 /* caller-supplied buffers (dogm_set_noise) */

typedef struct dogm_options
{
    uint64_t seed;     /* Philox key (default 123456, the reference's seed) */
    int resample_mode; /* DOGM_RESAMPLE_* */
    int noise_mode;    /* DOGM_NOISE_* */
} dogm_options;
int dogm_set_options(dogm_handle* h, const dogm_options* opts);
int dogm_get_options(const dogm_handle* h, dogm_options* out);

/* Injected noise for the NEXT stages that need it (device or host pointers; copied into handle-owned buffers).
 *   predict_noise : N * 4 floats, per particle (nx, ny, nvx, nvy) already scaled by the process-noise sigmas,
 *                   i.e. the values predictKernel adds (predict.cu:27-33)
 *   birth_noise   : B * 2 floats, per birth slot (vx, vy) already scaled by stddev_velocity (init_new_particles.cu:178-187)
 *   init_velocity : N * 2 floats, first-cycle velocities in [-init_max_velocity, +) (init_new_particles.cu:116-117)
 *   resample_unit : N floats, ascending, in [0,1): offsets as a fraction of the weight total
 * Any pointer may be NULL (that buffer keeps its previous content). */
int dogm_set_noise(dogm_handle* h, const float* predict_noise, const float* birth_noise, const float* init_velocity,
                   const float* resample_unit, int on_device);

/* Writes the noise the Philox stream of cycle `cycle` will produce into host buffers (same shapes as above;
 * resample_unit gets the systematic/stratified fractions (i + u)/N).  Lets a checker replay a Philox-mode cycle. */
int dogm_export_philox_noise(dogm_handle* h, uint32_t cycle, float* predict_noise, float* birth_noise,
                             float* init_velocity, float* resample_unit);
uint32_t dogm_get_cycle_counter(const dogm_handle* h);

/* State injection for stage-wise checks */
int dogm_set_particles(dogm_handle* h, const void* block, int on_device);              /* particle block of N */
int dogm_set_birth_particles(dogm_handle* h, const void* block, int on_device);        /* particle block of B */
int dogm_set_grid_cells(dogm_handle* h, const dogm_grid_cell* cells, int on_device);   /* grid_cell_count cells */
int dogm_set_measurement_cells(dogm_handle* h, const dogm_meas_cell* cells, int on_device);
int dogm_set_pose(dogm_handle* h, float x, float y, float yaw);                        /* as if a first pose had been received */
int dogm_get_birth_particles(dogm_handle* h, void* out_block);                         /* particle block of B */
int dogm_get_weight_array(dogm_handle* h, float* out_host);                            /* N */
int dogm_get_born_masses(dogm_handle* h, float* out_host);                             /* grid_cell_count */
int dogm_get_resampled_indices(dogm_handle* h, int* out_host);                         /* N */
int dogm_get_joint_weight_accum(dogm_handle* h, double* out_host);                     /* N + B */
int dogm_get_cell_ranges(dogm_handle* h, int* start_host, int* end_host);              /* grid_cell_count each */

/* The ancestor search of the resampling stage on caller-supplied data (resampling.cu:34-47):
 * out[i] = min(first j with cdf[j] >= draws[i], n_cdf - 1).  Host pointers; float CDF as in the reference. */
int dogm_search_ancestors_f32(dogm_handle* h, const float* cdf, int n_cdf, const float* sorted_draws, int n_draws,
                              int* out);

/* ------------------------------------------------------------------------------------------------------------
 * Measurement grid from a lidar scan (reference LaserMeasurementGrid, laser_to_meas_grid.h:13-35,
 * laser_to_meas_grid.cu:10-70, kernels measurement_grid.cu:33-89,115-132, GL warp renderer.cpp:11-31,66-84 +
 * shader.cpp:22-35 replaced by one CUDA kernel)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct dogm_laser_params /* == LaserMeasurementGrid::Params, laser_to_meas_grid.h:16-22 */
{
    float max_range;    /* [m] */
    float resolution;   /* [m/bin] */
    float fov;          /* [deg] */
    float stddev_range; /* [m] */
} dogm_laser_params;

typedef struct dogm_meas_handle dogm_meas_handle;

/* LaserMeasurementGrid(params, grid_length, resolution), laser_to_meas_grid.cu:10-18 */
int dogm_meas_create(const dogm_laser_params* params, float grid_length, float resolution, dogm_meas_handle** out);
void dogm_meas_destroy(dogm_meas_handle* m);
/* generateGrid(measurements), laser_to_meas_grid.cu:25-70: host beam ranges (inf = no return) -> device
 * MeasurementCell[grid_size^2] owned by the generator.  Synchronous like the reference (:67). */
int dogm_meas_generate(dogm_meas_handle* m, const float* beam_ranges_host, int num_beams, dogm_meas_cell** out_device);
/* Same scan for the next cycle of `h`, stream-ordered on the stream of `h`: the polar table of the scan is built now (up to 960
 * beams travel in the kernel parameters: no copy node), the cartesian resampling is done by the per-cell kernel of the following
 * dogm_update_grid(h, NULL, ...), which writes h's measurement buffer on the way - no measurement-grid pass and no 16*C-byte
 * copy (dogm.cu:207-208) of its own.  Whoever looks at the measurement buffer earlier (dogm_get_measurement_cells,
 * dogm_get_device_ptrs, a stage call, the first cycle's particle initialisation) gets it produced then.  `m` has to stay alive
 * until that update has been enqueued; a measurement grid passed to the update supersedes the scan.  Not for band handles. */
int dogm_meas_generate_into(dogm_meas_handle* m, dogm_handle* h, const float* beam_ranges_host, int num_beams);
/* Several scans of the same sensor geometry fused in the polar grid before the polar->cartesian step
 * (createPolarGridTextureKernel for the first scan, fusePolarGridTextureKernel + combine_masses for the others,
 * measurement_grid.cu:13-31,72-113; the reference ships the fusion kernel but never calls it).  scans_host is
 * [num_scans][num_beams]; out_polar_host (optional) receives the fused polar grid, [H][num_beams] pairs (occ, free). */
int dogm_meas_generate_fused(dogm_meas_handle* m, const float* scans_host, int num_scans, int num_beams,
                             dogm_meas_cell** out_device, float* out_polar_host);
/* The polar inverse-sensor-model grid alone (createPolarGridTextureKernel, measurement_grid.cu:72-89):
 * out_host[(range_bin * num_beams + beam) * 2 + {0: occ, 1: free}], range bins = int(max_range / resolution). */
int dogm_meas_polar_grid(dogm_meas_handle* m, const float* beam_ranges_host, int num_beams, float* out_host);
int dogm_meas_get_grid_size(const dogm_meas_handle* m);

/* ------------------------------------------------------------------------------------------------------------
 * On-device extraction of dynamic cells (reference computeCellsWithVelocity, demo/utils/image_creation.cpp:19-66):
 * cells with pignistic occupancy >= min_occupancy and Mahalanobis velocity v' S^-1 v >= min_velocity.
 * out_host receives up to `capacity` records; returns the number found through *out_count.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct dogm_dynamic_cell
{
    int cell_idx;
    float occupancy;
    float mean_x_vel;
    float mean_y_vel;
    float var_x_vel;
    float var_y_vel;
    float covar_xy_vel;
    float mahalanobis;
} dogm_dynamic_cell;
int dogm_extract_dynamic_cells(dogm_handle* h, float min_occupancy, float min_velocity, dogm_dynamic_cell* out_host,
                               int capacity, int* out_count);
/* Registers the filter ahead of time (capacity 0 switches it off): every following cycle then compacts the matching
 * cells inside its per-cell kernel, straight into pinned host memory, and dogm_extract_dynamic_cells with the same
 * thresholds returns that list instead of running a pass over all 64*C bytes of grid cells.  After dogm_update_grid_async it
 * returns as soon as the list exists - a kernel behind the per-cell kernel writes {count, sequence number} into host-mapped
 * memory the call spins on - while birth, CDF and resampling of that cycle are still running: the caller consumes the list and
 * enqueues the next scan in the meantime, so the stream never runs dry.  (Every other getter still synchronises.) */
int dogm_set_dynamic_cell_filter(dogm_handle* h, float min_occupancy, float min_velocity, int capacity);

/* ------------------------------------------------------------------------------------------------------------
 * Instrumentation for bench.py
 * ---------------------------------------------------------------------------------------------------------- */
void* dogm_get_stream(dogm_handle* h);                 /* cudaStream_t every kernel of this handle is launched on */
uint64_t dogm_get_launch_count(const dogm_handle* h);  /* kernels launched by this handle so far */
/* Per-kernel device timing with CUDA events on the handle's stream.  enable=1 starts recording an event pair
 * around every launch of subsequent cycles; dogm_kernel_timing_read sums them per kernel name since the last read. */
#define DOGM_MAX_KERNEL_SLOTS 48
typedef struct dogm_kernel_time
{
    char name[48];
    double total_ms;
    uint64_t launches;
    double algorithmic_bytes; /* per-launch algorithmic bytes of the last launch */
} dogm_kernel_time;
int dogm_kernel_timing_enable(dogm_handle* h, int enable);
/* Timeline of the free-running cycle without events between the kernels: while armed, the first CTA of every launch
 * stamps the GPU's nanosecond timer once its predecessor has completed; dogm_trace_read returns the stamps (48-byte
 * names, "#2" = second pass of the same kernel) of the launches since the last read and re-arms.  The difference of
 * two consecutive stamps is the earlier kernel's share of the cycle.  One handle per process at a time. */
int dogm_trace_arm(dogm_handle* h, int enable);
/* Developer read-out of an internal buffer by name ("res_start", "weight_total"); synchronises. */
int dogm_debug_read(dogm_handle* h, const char* name, void* out_host, size_t bytes);
int dogm_trace_read(dogm_handle* h, uint64_t* out_start_ns, char* out_names, int capacity, int* out_count);
int dogm_kernel_timing_read(dogm_handle* h, dogm_kernel_time* out, int capacity, int* out_count);

/* A CUDA-event stopwatch on the handle's stream (device time of everything enqueued between start and stop) */
int dogm_timer_start(dogm_handle* h);
int dogm_timer_stop(dogm_handle* h);
int dogm_timer_elapsed_ms(dogm_handle* h, float* out_ms); /* synchronises on the stop event */

/* Pinned host memory helpers so that callers without a CUDA runtime binding can stage inputs/outputs */
int dogm_host_alloc_pinned(void** out, size_t bytes);
int dogm_host_free_pinned(void* p);
int dogm_device_alloc(void** out, size_t bytes);
int dogm_device_free(void* p);
int dogm_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes);
int dogm_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes);
int dogm_memcpy_d2d(void* dst_device, const void* src_device, size_t bytes); /* same GPU or peer GPU of this process */
int dogm_device_count(void);
/* peer access from `device` to `peer_device` (NVLink / NVSwitch): device-to-device copies between them then bypass the host */
int dogm_enable_peer_access(int device, int peer_device);
int dogm_set_device(int device);
const char* dogm_b200_version(void);

#ifdef __cplusplus
}
#endif

#endif /* DOGM_B200_H */
