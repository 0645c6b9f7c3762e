"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs and identical
injected noise, free-running over several cycles.

Bar (BASELINE.json north_star): particle-to-cell indices and resample ancestor indices bit-exact; per-cell
occupied/free masses and mean velocities within 1e-4 relative.  Because every order-dependent sum on both sides is
accumulated in double and rounded once, the whole particle state is in fact expected to be bit-identical, which is
what these tests assert; velocity moments (float sums in a different order) are compared with a tolerance."""
import numpy as np
import pytest

from conftest import cycle_noise, make_params, synthetic_meas

pytestmark = pytest.mark.gpu

MASS_FIELDS = ["new_born_occ_mass", "pers_occ_mass", "free_mass", "occ_mass", "pred_occ_mass", "w_A", "w_UA"]
MOMENT_FIELDS = ["mean_x_vel", "mean_y_vel", "var_x_vel", "var_y_vel", "covar_xy_vel"]


def assert_particles_equal(got, exp, what, check_weight=True):
    assert np.array_equal(got.grid_cell_idx, exp.grid_cell_idx), f"{what}: cell indices differ"
    assert np.array_equal(got.state.view(np.uint32), exp.state.view(np.uint32)), f"{what}: states differ"
    if check_weight:
        assert np.array_equal(got.weight.view(np.uint32), exp.weight.view(np.uint32)), f"{what}: weights differ"
    assert np.array_equal(got.associated, exp.associated), f"{what}: associated flags differ"


def assert_cells_match(g, e, vel_scale):
    assert np.array_equal(g["start_idx"], e["start_idx"])
    assert np.array_equal(g["end_idx"], e["end_idx"])
    for f in MASS_FIELDS:
        # north_star tolerance is 1e-4 relative; the design gives bit-equality, so the check is much tighter
        assert np.allclose(g[f], e[f], rtol=1e-6, atol=1e-9), f
    occ = e["start_idx"] >= 0
    for f in ("mu_A", "mu_UA"):  # defined on non-empty cells only
        assert np.allclose(g[f][occ], e[f][occ], rtol=1e-6, atol=1e-9), f
    for f in MOMENT_FIELDS[:2]:
        assert np.allclose(g[f], e[f], rtol=1e-4, atol=1e-4 * vel_scale), f
    for f in MOMENT_FIELDS[2:]:
        assert np.allclose(g[f], e[f], rtol=1e-3, atol=1e-4 * vel_scale * vel_scale), f


def compare_after_cycle(d, o, p, tag, ties=0, draws=None):
    N = d.particle_count
    if ties:
        # Millions of weights spanning more than 2^29: the double sums of the two sides round differently in their last bits
        # (the comment at the CDF check below), and a draw that falls within that distance of a CDF entry picks the neighbouring
        # ancestor.  At most `ties` such draws are accepted, each one checked to be such a tie; the populations must agree
        # everywhere else, and the handle continues from the oracle's population so that the two stay in step.
        ga, oa = d.get_resampled_indices(), o.resampled_idx
        bad = np.flatnonzero(ga != oa)
        assert bad.size <= ties, f"{tag}: {bad.size} ancestors differ"
        gc, oc = d.get_joint_weight_accum(), o.joint_weight_accum
        for i in bad:
            lo, hi = min(ga[i], oa[i]), max(ga[i], oa[i])
            assert hi - lo == 1, f"{tag}: slot {i}: ancestors {ga[i]} / {oa[i]} are not neighbours"
            r = draws(i, gc[-1])
            assert abs(gc[lo] - r) <= 1e-10 * abs(r) and abs(oc[lo] - r) <= 1e-10 * abs(r), f"{tag}: slot {i} is no tie"
        gp, op = d.get_particles(), o.particles
        keep = np.ones(N, bool)
        keep[bad] = False
        assert np.array_equal(gp.grid_cell_idx[keep], op.grid_cell_idx[keep]) and np.array_equal(gp.associated[keep], op.associated[keep])
        assert np.array_equal(gp.state.view(np.uint32)[keep], op.state.view(np.uint32)[keep]), f"{tag}: states differ"
        assert np.array_equal(gp.weight.view(np.uint32), op.weight.view(np.uint32)), f"{tag}: weights differ"
        if bad.size:
            d.set_particles(type(gp).from_arrays(op.state, op.grid_cell_idx, op.weight, op.associated))
    else:
        assert_particles_equal(d.get_particles(), o.particles, f"{tag} particles")
    assert_particles_equal(d.get_birth_particles(), o.birth_particles, f"{tag} birth particles")
    if not ties:
        assert np.array_equal(d.get_resampled_indices(), o.resampled_idx), f"{tag}: ancestor indices differ"
    assert np.array_equal(d.get_weight_array().view(np.uint32), o.weight_array.view(np.uint32)), f"{tag}: weight_array"
    assert np.array_equal(d.get_born_masses().view(np.uint32), o.born_masses.view(np.uint32)), f"{tag}: born masses"
    # Both sides add the same float weights in double, the oracle one after the other, the kernel tile-wise.  While the weights
    # span less than 2^29 the partial sums are exact in either order (24-bit addends, 53-bit accumulator) and the arrays are
    # equal to the last bit; with very small weights in the set (a birth weight of 1e-12 next to persistent weights of 1e-4) the
    # two orders round differently in the last bits: a serial double sum of k terms is off by about sqrt(k) * 2^-53 relative,
    # 1e-13 at k = 3e5.  What decides the ancestors is checked bit for bit above.
    assert np.allclose(d.get_joint_weight_accum(), o.joint_weight_accum, rtol=1e-11 if ties else 1e-12, atol=0), f"{tag}: cdf"
    assert_cells_match(d.get_grid_cells(), o.grid_cells, p.stddev_velocity)
    assert (d.get_position_x(), d.get_position_y(), d.get_yaw()) == o.position


def free_run(gpu, orc, size, res, n, b, cycles, seed, ego=(0.0, 0.45), dt=0.1, meas_every=1, systematic=False, ties=0, free_level=0.3,
             **over):
    rng = np.random.default_rng(seed)
    p = make_params(gpu, size, res, n, b, **over)
    po = make_params(orc, size, res, n, b, **over)
    d = gpu.DOGM(p)
    d.set_options(noise_mode=gpu.NOISE_INJECTED, resample_mode=gpu.RESAMPLE_SYSTEMATIC if systematic else gpu.RESAMPLE_INJECTED)
    o = orc.OracleDOGM(po, resample_mode=orc.RESAMPLE_SYSTEMATIC if systematic else orc.RESAMPLE_INJECTED)
    gs = d.grid_size
    assert gs == o.grid_size
    meas = None
    for c in range(cycles):
        if c % meas_every == 0:
            meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, gs, rng, free_level=free_level)
        pn, bn, iv, ru = cycle_noise(rng, n, b, p)
        d.set_noise(pn, bn, iv, ru)
        o.set_noise(pn, bn, iv, ru)
        x, y = ego[0] * c, ego[1] * c
        d.update_grid(meas, x, y, 0.0, dt, device=False)
        o.update_grid(meas.view(orc.MEAS_CELL_DTYPE), x, y, 0.0, dt)
        if systematic:  # the offset of output slot i: (i + u) * total / N in double (one fraction for all slots)
            draws = lambda i, total, u=float(ru[0]): (float(i) + u) * (total / n)
        else:  # injected: float(total) * u_i in float, resampling.cu:34-47
            draws = lambda i, total, ru=ru: float(np.float32(total) * ru[i])
        compare_after_cycle(d, o, p, f"cycle {c}", ties=ties, draws=draws)
    d.close()
    o.close()


def test_free_running_single_pass_sort(gpu, orc):
    # 40x40 cells -> 11 key bits -> one counting-sort pass; ego motion shifts the grid on most cycles
    free_run(gpu, orc, 10.0, 0.25, 20000, 2000, cycles=6, seed=11)


def test_free_running_two_pass_sort(gpu, orc):
    # 128x128 cells -> 14 key bits -> two passes of 7 bits
    free_run(gpu, orc, 64.0, 0.5, 100000, 10000, cycles=5, seed=12, ego=(0.3, 0.8))


@pytest.mark.parametrize("size,res,n,b,ego", [(64.0, 0.5, 100000, 10000, (0.3, 0.8)), (50.0, 0.2, 300000, 30000, (0.0, 0.4)),
                                              (120.0, 0.25, 37000, 900, (-0.6, 0.2))])
def test_free_running_bucket_sort(gpu, orc, monkeypatch, size, res, n, b, ego):
    """The optional bucket sort (DOGM_B200_SORT=bucket: splitters from the population's own cell order, grouping pass by bucket,
    one counting sort per bucket) against the oracle over free-running cycles, shifts included: same particle order, ranges,
    ancestors and masses as the radix passes give."""
    monkeypatch.setenv("DOGM_B200_SORT", "bucket")
    free_run(gpu, orc, size, res, n, b, cycles=5, seed=21, ego=ego)


@pytest.mark.parametrize("size,res,n,b,ego,over", [
    (64.0, 0.5, 100000, 10000, (0.3, 0.8), {}),
    (50.0, 0.2, 300000, 30000, (0.0, 0.4), {}),            # the demo size
    (120.0, 0.25, 37000, 900, (-0.6, 0.2), {}),            # few birth particles, ragged last tile
    (9.0, 0.3, 1001, 77, (-0.7, 0.35), {}),                # not a multiple of anything
    (10.0, 1.0, 2, 1, (1.5, -2.5), {}),                    # two particles
    (33.0, 0.5, 4097, 513, (0.0, 0.0), {}),
    (8.0, 0.5, 3000, 0, (0.0, 0.6), {}),                   # no birth particles at all
    (4.0, 0.5, 30000, 3000, (0.0, 0.0), dict(stddev_process_noise_position=0.01, stddev_process_noise_velocity=0.1,
                                             stddev_velocity=0.5, init_max_velocity=0.5)),  # 64 cells: long runs of copies
])
def test_free_running_systematic_resampling(gpu, orc, size, res, n, b, ego, over):
    """Systematic resampling (BASELINE.json's production mode: one fraction for all offsets; window starts claimed by the CDF
    kernel, search in shared memory) against the oracle's lower_bound on the same CDF with the same fraction: ancestors, the
    next population and everything downstream bit for bit, over free-running cycles with shifts - ragged sizes, two particles,
    no birth particles, long runs of copies of one ancestor."""
    free_run(gpu, orc, size, res, n, b, cycles=5, seed=41, ego=ego, systematic=True, **over)


@pytest.mark.parametrize("size,res,n,b,ego", [(64.0, 0.5, 100000, 10000, (0.3, 0.8)), (10.0, 0.25, 20000, 2000, (0.0, 0.45)),
                                              (50.0, 0.2, 300000, 30000, (0.0, 0.4))])
def test_free_running_separate_scan_kernels_at_small_sizes(gpu, orc, monkeypatch, size, res, n, b, ego):
    """Up to 128 sort tiles the scatter kernels scan the tile histograms themselves and count the next pass's digits while they
    store (k_scatter FUSE: three launches fewer per cycle - what every small case of this file runs).  DOGM_B200_SORT=radix keeps
    the separate kernels the large sizes use (k_hist_scan, k_pair_tile_hist) at these sizes too: same results."""
    monkeypatch.setenv("DOGM_B200_SORT", "radix")
    free_run(gpu, orc, size, res, n, b, cycles=4, seed=51, ego=ego)


def test_free_running_three_pass_sort(gpu, orc, monkeypatch):
    """Cell keys of more than 24 bits (the 16384^2 grid: 28) are sorted in three passes; with the digit width limited to 5 bits
    (DOGM_B200_MAX_DIGIT_BITS) a 128 x 128 grid (14 key bits) takes the same plan: same order, ranges, ancestors as the oracle."""
    monkeypatch.setenv("DOGM_B200_MAX_DIGIT_BITS", "5")
    free_run(gpu, orc, 64.0, 0.5, 100000, 10000, cycles=4, seed=19, ego=(0.3, 0.8))
    free_run(gpu, orc, 64.0, 0.5, 4097, 513, cycles=3, seed=20, ego=(-0.5, 0.25), systematic=True)


@pytest.mark.parametrize("systematic", [False, True])
def test_free_running_more_cdf_tiles_than_resident_ctas(gpu, orc, systematic):
    """4.4e6 joint-weight entries = 1075 CDF tiles: more than the chained scan keeps resident at once (its CTAs walk over the
    tickets), more than the flat look-back takes (groups of 256 tiles + group prefixes), and no window starts for the
    resampling CTAs (they search) - the paths the 4096^2 and 16384^2 grids run, here against the oracle; 260 particles per cell
    on average, so most cells span several chunks of the segmented reduction."""
    free_run(gpu, orc, 64.0, 0.5, 4_300_000, 100_000, cycles=3, seed=23, ego=(0.3, 0.8), systematic=systematic, ties=4)


def test_free_running_4096_grid_12_bit_digits_and_quiet_blocks(gpu, orc):
    """4096 x 4096 cells (BASELINE.json configs[2]'s grid) with few particles: 24 key bits = two sort passes of 4096 bins, the
    widest digits there are, and from the third cycle on the cell kernel's shortcut for blocks in which nothing happens - both
    against the oracle, cell by cell (the property tests of test_gpu_fullsize.py run this grid with its 2.2e7 particles)."""
    # (no free mass outside the occupied blobs: most blocks of 256 cells see no particles, no measurement and no free mass)
    free_run(gpu, orc, 409.6, 0.1, 200_000, 20_000, cycles=5, seed=29, ego=(0.0, 0.0), meas_every=5, free_level=0.0)
    free_run(gpu, orc, 409.6, 0.1, 200_000, 20_000, cycles=2, seed=30, ego=(0.0, 0.45))


def test_free_running_config1_reference_demo(gpu, orc):
    # BASELINE.json configs[0]: 250x250 grid, 3e5 persistent + 3e4 birth particles (README.md:20-22)
    free_run(gpu, orc, 50.0, 0.2, 300000, 30000, cycles=4, seed=13, ego=(0.0, 0.4))


def test_free_running_ragged_sizes(gpu, orc):
    # particle counts that are not multiples of 4 / 32 / the tile sizes; x and y shifts of both signs
    free_run(gpu, orc, 9.0, 0.3, 1001, 77, cycles=5, seed=14, ego=(-0.7, 0.35))
    free_run(gpu, orc, 10.0, 1.0, 2, 1, cycles=4, seed=15, ego=(1.5, -2.5))
    free_run(gpu, orc, 33.0, 0.5, 4097, 513, cycles=3, seed=16, ego=(0.0, 0.0))


def test_free_running_dense_scene_over_unit_cells(gpu, orc):
    # few cells, many particles, no process noise on position: predicted cell masses exceed 1 and get renormalised
    free_run(gpu, orc, 4.0, 0.5, 30000, 3000, cycles=4, seed=17, ego=(0.0, 0.0),
             stddev_process_noise_position=0.01, stddev_process_noise_velocity=0.1, stddev_velocity=0.5, init_max_velocity=0.5)


def test_null_measurement_grid(gpu, orc):
    """updateGrid(nullptr, ...) (dogm_spec.cpp:32): the initial measurement grid stays; the cycle still runs."""
    p = make_params(gpu, 10.0, 1.0, 64, 8)
    po = make_params(orc, 10.0, 1.0, 64, 8)
    d = gpu.DOGM(p)
    d.set_options(noise_mode=gpu.NOISE_INJECTED, resample_mode=gpu.RESAMPLE_INJECTED)
    o = orc.OracleDOGM(po)
    rng = np.random.default_rng(18)
    for c in range(3):
        pn, bn, iv, ru = cycle_noise(rng, 64, 8, p)
        d.set_noise(pn, bn, iv, ru)
        o.set_noise(pn, bn, iv, ru)
        d.update_grid(None, 1.0 * c, 0.0, 0.0, 0.1)
        o.update_grid(None, 1.0 * c, 0.0, 0.0, 0.1)
        compare_after_cycle(d, o, p, f"null-meas cycle {c}")
    m = d.get_measurement_cells()
    assert np.all(m["likelihood"] == 1.0) and np.all(m["p_A"] == 1.0) and np.all(m["occ_mass"] == 0.0)


def test_stage_api_equals_update_grid(gpu, orc):
    """The eight public stage methods (dogm.h:148-156) run one after the other give the same cycle as updateGrid."""
    rng = np.random.default_rng(19)
    n, b = 30000, 3000
    p = make_params(gpu, 20.0, 0.5, n, b)
    a, s = gpu.DOGM(p), gpu.DOGM(p)
    for d in (a, s):
        d.set_options(noise_mode=gpu.NOISE_INJECTED, resample_mode=gpu.RESAMPLE_INJECTED)
    meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, a.grid_size, rng)
    for c in range(3):
        pn, bn, iv, ru = cycle_noise(rng, n, b, p)
        a.set_noise(pn, bn, iv, ru)
        s.set_noise(pn, bn, iv, ru)
        a.update_grid(meas, 0.0, 0.0, 0.0, 0.1, device=False)
        s.set_measurement_cells(meas)
        if c == 0:
            s.initialize_particles()
        s.particle_prediction(0.1)
        s.particle_assignment()
        s.grid_cell_occupancy_update(0.1)
        s.update_persistent_particles()
        s.initialize_new_particles()
        s.statistical_moments()
        s.resampling()
        assert_particles_equal(s.get_particles(), a.get_particles(), f"stage-wise cycle {c}")
        ga, gs_ = a.get_grid_cells(), s.get_grid_cells()
        assert np.array_equal(ga.view(np.uint8), gs_.view(np.uint8))


def test_spec_predict_on_gpu(gpu):
    """DOGM.Predict, test/dogm_spec.cpp:68-103, against the CUDA library."""
    p = make_params(gpu, 10.0, 1.0, 2, 1, persistence_prob=0.5, stddev_process_noise_position=0.0,
                    stddev_process_noise_velocity=0.0, stddev_velocity=10.0)
    d = gpu.DOGM(p)
    parts = gpu.ParticlesSoA.from_arrays([[3.25, 4.5, 1.7, -2.3], [7.125, 1.0625, -30.0, 12.5]], [0, 0], [0.37, 0.63])
    d.set_particles(parts)
    dt = np.float32(0.1)
    d.particle_prediction(float(dt))  # Philox noise with sigma = 0 adds exactly 0
    q = d.get_particles()
    pred = parts.state.copy()
    pred[:, 0] = parts.state[:, 0] + dt * parts.state[:, 2]
    pred[:, 1] = parts.state[:, 1] + dt * parts.state[:, 3]
    assert np.array_equal(q.state, pred)
    assert np.array_equal(q.weight, parts.weight * np.float32(0.5))


def test_spec_ego_motion_compensation_on_gpu(gpu):
    """DOGM.EgoMotionCompensation, test/dogm_spec.cpp:10-66: pose bookkeeping and the +3 cell particle shift."""
    p = make_params(gpu, 10.0, 1.0, 2, 1, persistence_prob=0.5, stddev_process_noise_position=0.0,
                    stddev_process_noise_velocity=0.0, stddev_velocity=10.0)
    d = gpu.DOGM(p)
    d.update_grid(None, 10.0, 10.0, 0.0, 0.0)
    assert (d.get_position_x(), d.get_position_y()) == (10.0, 10.0)
    d.update_grid(None, 10.5, 10.5, 0.0, 0.0)
    assert (d.get_position_x(), d.get_position_y()) == (10.0, 10.0)
    parts = gpu.ParticlesSoA.from_arrays([[3.25, 4.5, 0.0, 0.0], [5.5, 5.5, 0.0, 0.0]], [0, 0], [0.5, 0.5])
    d.set_particles(parts)
    d.update_grid(None, 13.0, 10.0, 0.0, 0.0)
    assert (d.get_position_x(), d.get_position_y()) == (13.0, 10.0)
    q = d.get_particles()
    allowed = {(3.25 + 3.0, 4.5), (5.5 + 3.0, 5.5), (0.5, 0.5)}  # shifted survivors, or a newborn at cell 0's centre
    for s in q.state:
        assert (float(s[0]), float(s[1])) in allowed


def philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def test_philox_known_answers_and_device_stream(gpu):
    # Random123 known-answer vectors for philox4x32-10
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    # the device stream: resampling fractions are u = (word0 >> 8) * 2^-24 of counter (slot, STAGE_RESAMPLE=4, cycle, 0)
    p = make_params(gpu, 10.0, 1.0, 64, 8)
    d = gpu.DOGM(p)
    seed = 0x1234ABCD5678
    d.set_options(seed=seed, resample_mode=gpu.RESAMPLE_STRATIFIED)
    _, _, _, ru = d.export_philox_noise(7)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    exp = np.array([(philox4x32_10((i, 4, 7, 0), key)[0] >> 8) / 2.0**24 for i in range(64)], np.float32)
    assert np.array_equal(ru, exp)
    pn, bn, iv, _ = d.export_philox_noise(7)
    assert np.all(np.isfinite(pn)) and np.all(np.abs(iv) <= 30.0)


@pytest.mark.parametrize("mode", ["systematic", "stratified"])
def test_philox_mode_replayed_by_oracle(gpu, orc, mode):
    """Production mode: Philox noise generated on the device; the same numbers exported and fed to the oracle."""
    rng = np.random.default_rng(21)
    n, b = 50000, 5000
    p = make_params(gpu, 25.0, 0.5, n, b)
    po = make_params(orc, 25.0, 0.5, n, b)
    gmode = gpu.RESAMPLE_SYSTEMATIC if mode == "systematic" else gpu.RESAMPLE_STRATIFIED
    omode = orc.RESAMPLE_SYSTEMATIC if mode == "systematic" else orc.RESAMPLE_STRATIFIED
    d = gpu.DOGM(p)
    d.set_options(seed=987654321, resample_mode=gmode, noise_mode=gpu.NOISE_PHILOX)
    o = orc.OracleDOGM(po, resample_mode=omode)
    for c in range(4):
        meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, d.grid_size, rng)
        pn, bn, iv, ru = d.export_philox_noise(d.cycle_counter())
        o.set_noise(pn, bn, iv, ru)
        d.update_grid(meas, 0.0, 0.45 * c, 0.0, 0.1, device=False)
        o.update_grid(meas.view(orc.MEAS_CELL_DTYPE), 0.0, 0.45 * c, 0.0, 0.1)
        compare_after_cycle(d, o, p, f"philox {mode} cycle {c}")
    # determinism: a second handle with the same seed reproduces the population bit for bit
    d2 = gpu.DOGM(p)
    d2.set_options(seed=987654321, resample_mode=gmode, noise_mode=gpu.NOISE_PHILOX)
    rng = np.random.default_rng(21)
    for c in range(4):
        meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, d2.grid_size, rng)
        d2.update_grid(meas, 0.0, 0.45 * c, 0.0, 0.1, device=False)
    assert np.array_equal(d2.get_particles().block, d.get_particles().block)


def test_search_ancestors_f32_kernel(gpu, orc):
    """resampling.cu:34-47 on a float CDF (the reference's data type), incl. zero-weight entries and ties."""
    rng = np.random.default_rng(22)
    w = rng.uniform(0, 1, 200000).astype(np.float32)
    w[rng.integers(0, w.size, 20000)] = 0.0
    cdf = np.cumsum(w, dtype=np.float32)
    draws = np.sort(rng.uniform(0, cdf[-1], 150000).astype(np.float32))
    draws[:10] = 0.0
    draws[-1] = cdf[-1]
    p = make_params(gpu, 10.0, 1.0, 4, 1)
    d = gpu.DOGM(p)
    got = d.search_ancestors_f32(cdf, draws)
    assert np.array_equal(got, orc.search_ancestors_f32(cdf, draws))
    assert np.array_equal(got, np.minimum(np.searchsorted(cdf, draws, side="left"), cdf.size - 1))


def test_dynamic_cell_extraction(gpu, orc):
    rng = np.random.default_rng(23)
    n, b = 60000, 6000
    p = make_params(gpu, 20.0, 0.5, n, b)
    d = gpu.DOGM(p)
    for c in range(6):
        meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, d.grid_size, np.random.default_rng(5))
        d.update_grid(meas, 0.0, 0.0, 0.0, 0.1, device=False)
    cells = d.get_grid_cells()
    got, count = d.extract_dynamic_cells(0.6, 0.5)
    rec, n_exp = orc.extract_dynamic_cells(cells.view(orc.GRID_CELL_DTYPE), 0.6, 0.5)
    assert count == n_exp
    order = np.argsort(got["cell_idx"])
    exp_idx = rec[:, 0].copy().view(np.int32)
    assert np.array_equal(got["cell_idx"][order], np.sort(exp_idx))
    eo = np.argsort(exp_idx)
    assert np.allclose(got["occupancy"][order], rec[eo, 1], rtol=1e-6)
    assert np.allclose(got["mahalanobis"][order], rec[eo, 7], rtol=1e-4, atol=1e-5)


def test_dynamic_cell_filter_fused_into_cycle(gpu, orc):
    """dogm_set_dynamic_cell_filter: the list compacted by the cycle's own cell kernel equals the stand-alone pass over
    the published grid cells (computeCellsWithVelocity, demo/utils/image_creation.cpp:19-66), with ego-motion shifts
    between the cycles; other thresholds fall back to the stand-alone pass; a too small capacity truncates only."""
    n, b = 60000, 6000
    p = make_params(gpu, 20.0, 0.5, n, b)
    d = gpu.DOGM(p)
    d.set_dynamic_cell_filter(0.6, 0.5, 4096)
    for c in range(6):
        meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, d.grid_size, np.random.default_rng(5 + c))
        d.update_grid(meas, 0.7 * c, -0.4 * c, 0.0, 0.1, device=False)
        got, count = d.extract_dynamic_cells(0.6, 0.5, capacity=4096)  # fast path
        cells = d.get_grid_cells()
        rec, n_exp = orc.extract_dynamic_cells(cells.view(orc.GRID_CELL_DTYPE), 0.6, 0.5)
        assert count == n_exp
        order = np.argsort(got["cell_idx"])
        exp_idx = rec[:, 0].copy().view(np.int32)
        eo = np.argsort(exp_idx)
        assert np.array_equal(got["cell_idx"][order], exp_idx[eo])
        assert np.allclose(got["occupancy"][order], rec[eo, 1], rtol=1e-6)
        assert np.allclose(got["mean_x_vel"][order], rec[eo, 2], rtol=1e-6, atol=1e-7)
        assert np.allclose(got["mahalanobis"][order], rec[eo, 7], rtol=1e-4, atol=1e-5)
    assert count > 0
    # different thresholds: stand-alone pass, same answer as the oracle
    got2, count2 = d.extract_dynamic_cells(0.5, 0.25)
    rec2, n2 = orc.extract_dynamic_cells(cells.view(orc.GRID_CELL_DTYPE), 0.5, 0.25)
    assert count2 == n2 and np.array_equal(np.sort(got2["cell_idx"]), np.sort(rec2[:, 0].copy().view(np.int32)))
    # truncation: the count is the number found, the records are a subset
    small = max(1, count // 2)
    got3, count3 = d.extract_dynamic_cells(0.6, 0.5, capacity=small)
    assert count3 == count and len(got3) == small and set(got3["cell_idx"]) <= set(exp_idx.tolist())
    # switching the filter off restores the stand-alone behaviour
    d.set_dynamic_cell_filter(0.0, 0.0, 0)
    d.update_grid(meas, 4.0, -2.0, 0.0, 0.1, device=False)
    got4, count4 = d.extract_dynamic_cells(0.6, 0.5)
    rec4, n4 = orc.extract_dynamic_cells(d.get_grid_cells().view(orc.GRID_CELL_DTYPE), 0.6, 0.5)
    assert count4 == n4


def test_dynamic_cell_list_published_before_cycle_end(gpu, orc):
    """With a registered filter dogm_extract_dynamic_cells does not wait for the end of the cycle: the kernel behind the cell
    kernel publishes {count, sequence number} into host-mapped memory.  Asynchronous cycles (device-resident measurement
    grids, the next cycle enqueued straight after the read-out) give the same lists as the stand-alone pass over the
    published grid cells, and the final state equals that of a handle driven with blocking calls."""
    n, b = 200000, 20000
    p = make_params(gpu, 40.0, 0.2, n, b)
    d = gpu.DOGM(p)
    ref = gpu.DOGM(p)
    d.set_dynamic_cell_filter(0.6, 0.5, 8192)
    grids = []
    for c in range(5):
        meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, d.grid_size, np.random.default_rng(50 + c))
        ptr = gpu.device_alloc(meas.nbytes)
        gpu.memcpy_h2d(ptr, meas)
        grids.append(ptr)
    lists = []
    for c in range(5):
        d.update_grid(grids[c], 0.5 * c, 0.3 * c, 0.0, 0.1, device=True, sync=False)
        got, count = d.extract_dynamic_cells(0.6, 0.5, capacity=8192)  # returns while the cycle is still running
        lists.append((np.sort(got["cell_idx"]).copy(), count))
    for c in range(5):
        ref.update_grid(grids[c], 0.5 * c, 0.3 * c, 0.0, 0.1, device=True)
        rec, n_exp = orc.extract_dynamic_cells(ref.get_grid_cells().view(orc.GRID_CELL_DTYPE), 0.6, 0.5)
        assert lists[c][1] == n_exp
        assert np.array_equal(lists[c][0], np.sort(rec[:, 0].copy().view(np.int32)))
    assert lists[-1][1] > 0
    assert np.array_equal(d.get_grid_cells(), ref.get_grid_cells())
    assert np.array_equal(d.get_particles().block, ref.get_particles().block)
    # the stage API stops after the occupancy update: nothing publishes the list, the read-out falls back to a stream sync
    d.update_pose(3.0, 2.0, 0.0)
    d.particle_prediction(0.1)
    d.particle_assignment()
    d.grid_cell_occupancy_update(0.1)
    got, count = d.extract_dynamic_cells(0.6, 0.5, capacity=8192)
    rec, n_exp = orc.extract_dynamic_cells(d.get_grid_cells().view(orc.GRID_CELL_DTYPE), 0.6, 0.5)
    assert count == n_exp and np.array_equal(np.sort(got["cell_idx"]), np.sort(rec[:, 0].copy().view(np.int32)))
    for ptr in grids:
        gpu.device_free(ptr)


def test_concurrent_handles_do_not_interfere(gpu):
    """Independent sensor streams on one GPU (BASELINE.json configs[3]): handles driven interleaved and asynchronously on
    their own CUDA streams end in exactly the state each reaches when run alone (the kernels' inter-CTA waits - chained
    scan, bin-base chain - only ever wait on work of the same launch)."""
    size, res, n, b = 50.0, 0.2, 300_000, 30_000
    p = make_params(gpu, size, res, n, b)
    rngs = [np.random.default_rng(100 + k) for k in range(3)]
    meas = [[synthetic_meas(gpu.MEAS_CELL_DTYPE, 250, rngs[k]) for _ in range(3)] for k in range(3)]

    def fresh(k):
        d = gpu.DOGM(p)
        d.set_options(seed=900 + k, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
        return d

    alone = []
    for k in range(3):
        d = fresh(k)
        for c in range(8):
            d.update_grid(meas[k][c % 3], 0.1 * c, 0.4 * c, 0.0, 0.1, device=False)
        alone.append((d.get_particles().block.copy(), d.get_grid_cells().copy()))
        d.close()
    group = [fresh(k) for k in range(3)]
    dev = [[gpu.device_alloc(m.nbytes) for m in meas[k]] for k in range(3)]
    for k in range(3):
        for ptr, m in zip(dev[k], meas[k]):
            gpu.memcpy_h2d(ptr, m)
    for c in range(8):
        for k, d in enumerate(group):
            d.update_grid(dev[k][c % 3], 0.1 * c, 0.4 * c, 0.0, 0.1, device=True, sync=False)
    for k, d in enumerate(group):
        d.synchronize()
        assert np.array_equal(d.get_particles().block, alone[k][0]), f"handle {k}: particles differ"
        assert np.array_equal(d.get_grid_cells().view(np.uint8), alone[k][1].view(np.uint8)), f"handle {k}: cells differ"
        d.close()
    for row in dev:
        for ptr in row:
            gpu.device_free(ptr)


@pytest.mark.timeout(120)
def test_degenerate_inputs_terminate(gpu, orc):
    """Inputs the demo never produces must not hang or crash the cycle (every inter-CTA wait inside the kernels is on
    words that are published unconditionally): an all-zero measurement grid (no born mass: the birth normaliser is 0),
    a grid without any free or occupied evidence after the particles have died out, a huge time step that throws every
    particle out of the grid (total weight 0), and NaN / negative ranges in a scan."""
    n, b = 40_000, 4_000
    p = make_params(gpu, 20.0, 0.25, n, b)
    gs = 80
    zero = np.zeros(gs * gs, gpu.MEAS_CELL_DTYPE)
    zero["likelihood"] = 1.0
    zero["p_A"] = 1.0
    d = gpu.DOGM(p)
    for c in range(4):
        d.update_grid(zero, 0.0, 0.3 * c, 0.0, 0.1, device=False)
    cells = d.get_grid_cells()
    assert np.all(cells["occ_mass"][np.isfinite(cells["occ_mass"])] <= 1.0 + 1e-5)
    d.close()

    d = gpu.DOGM(p)
    rng = np.random.default_rng(8)
    meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, gs, rng)
    for c in range(3):
        d.update_grid(meas, 0.0, 0.0, 0.0, 0.1, device=False)
    d.update_grid(meas, 0.0, 0.0, 0.0, 1.0e6, device=False)  # every persistent particle leaves the grid
    d.update_grid(meas, 0.0, 0.0, 0.0, 0.1, device=False)
    pa = d.get_particles()
    assert np.all(np.isfinite(pa.weight)) and np.all(pa.grid_cell_idx >= 0) and np.all(pa.grid_cell_idx < gs * gs)
    for c in range(3):  # and the filter recovers from the birth particles
        d.update_grid(meas, 0.0, 0.0, 0.0, 0.1, device=False)
    assert float(d.get_grid_cells()["occ_mass"].sum()) > 1.0
    d.close()

    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(20.0, 0.25, 120.0, 0.5), 20.0, 0.25)
    z = np.array([np.nan, -1.0, 0.0, np.inf, 5.0, 1e9, -np.inf] * 6, np.float32)
    grid = gen.generate_grid_host(z)
    assert np.all(np.isfinite(grid["occ_mass"])) and np.all(np.isfinite(grid["free_mass"]))
    d = gpu.DOGM(p)
    for c in range(3):
        d.update_grid(grid, 0.0, 0.0, 0.0, 0.1, device=False)
    d.close()
    gen.close()


def test_pipelined_grid_cell_readback_equals_the_blocking_getter(gpu):
    """dogm_get_grid_cells_begin / _wait (the copy of cycle n's cells runs under cycle n + 1, the cell kernel alternates between
    two GridCell buffers): every cycle's cells arrive bit-identical to what dogm_get_grid_cells returns for a twin handle that
    is driven with blocking calls; the device pointer of the grid cells follows the buffer the last cycle wrote; the blocking
    getter still works in between; a read-out begun twice for one cycle delivers the same cells twice."""
    size, res, n, b = 64.0, 0.5, 60000, 6000
    p = make_params(gpu, size, res, n, b)
    twins = []
    for _ in range(2):
        d = gpu.DOGM(p)
        d.set_options(seed=99, noise_mode=gpu.NOISE_PHILOX, resample_mode=gpu.RESAMPLE_SYSTEMATIC)
        twins.append(d)
    a, pl = twins
    gs, C = a.grid_size, a.grid_cell_count
    rng = np.random.default_rng(5)
    meas = [gpu.pinned_empty((C,), gpu.MEAS_CELL_DTYPE) for _ in range(3)]
    for m in meas:
        m[:] = synthetic_meas(gpu.MEAS_CELL_DTYPE, gs, rng)
    bufs = [gpu.pinned_empty((C,), gpu.GRID_CELL_DTYPE) for _ in range(2)]
    extra = gpu.pinned_empty((C,), gpu.GRID_CELL_DTYPE)
    expected, ptrs = [], set()
    cycles = 7
    for c in range(cycles):
        x, y = 0.3 * c, 0.7 * c
        a.update_grid(meas[c % 3], x, y, 0.0, 0.1, device=False)
        expected.append(a.get_grid_cells().copy())
        pl.update_grid(meas[c % 3], x, y, 0.0, 0.1, device=False, sync=False)
        if c > 0:
            pl.get_grid_cells_wait()
            assert np.array_equal(bufs[(c - 1) & 1].view(np.uint8), expected[c - 1].view(np.uint8)), f"cycle {c - 1}"
        pl.get_grid_cells_begin(bufs[c & 1])
        if c == 3:
            pl.get_grid_cells_begin(extra)  # the same cycle once more
            assert np.array_equal(pl.get_grid_cells().view(np.uint8), expected[c].view(np.uint8))  # and through the blocking getter
        ptrs.add(int(pl.device_ptrs().grid_cell_array))
    pl.get_grid_cells_wait()
    assert np.array_equal(bufs[(cycles - 1) & 1].view(np.uint8), expected[-1].view(np.uint8))
    assert np.array_equal(extra.view(np.uint8), expected[3].view(np.uint8))
    assert len(ptrs) == 2  # two buffers in turn
    assert_particles_equal(pl.get_particles(), a.get_particles(), "populations of the twins")
    for d in twins:
        d.close()
