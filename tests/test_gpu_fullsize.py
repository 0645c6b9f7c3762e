"""Size-independent properties at BASELINE.json's headline size (1200 x 1200 cells, 2e6 + 2e5 particles), production
mode (Philox noise, systematic resampling), where the CPU oracle would take too long for a multi-cycle comparison:
sortedness, range consistency, mass conservation, ancestor = lower_bound on the device's own CDF, gather correctness,
determinism.  One extra cycle IS compared against the oracle with the exported Philox noise."""
import numpy as np
import pytest

from conftest import make_params

pytestmark = pytest.mark.gpu

SIZE, RES, N, B = 120.0, 0.1, 2_000_000, 200_000


def scene_meas(gpu, rng, beams=480, size=SIZE, res=RES):
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(size, res, 120.0, 0.5), size, res)
    z = np.full(beams, np.inf, np.float32)
    hit = rng.uniform(size=beams) < 0.55
    z[hit] = rng.uniform(5.0, 0.9 * size, size=int(hit.sum())).astype(np.float32)
    meas = gen.generate_grid_host(z)
    gen.close()
    return meas


# the second case is BASELINE.json's 4096 x 4096 grid with 2e7 + 2e6 particles on one GPU: 24-bit cell keys (three sort
# passes), more CDF tiles than the device holds at once (no window starts for the resampling CTAs: they search)
@pytest.mark.parametrize("SIZE,RES,N,B,beams,gs_expected", [(120.0, 0.1, 2_000_000, 200_000, 480, 1200),
                                                            (409.6, 0.1, 20_000_000, 2_000_000, 1638, 4096)])
def test_full_size_properties_and_determinism(gpu, orc, SIZE, RES, N, B, beams, gs_expected):
    rng = np.random.default_rng(51)
    p = make_params(gpu, SIZE, RES, N, B)
    d = gpu.DOGM(p)
    d.set_options(seed=4242, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
    gs, C = d.grid_size, d.grid_cell_count
    assert gs == gs_expected
    meas = scene_meas(gpu, rng, beams, SIZE, RES)
    for c in range(4):
        d.update_grid(meas, 0.0, 0.4 * c, 0.0, 0.1, device=False)

    # --- one more cycle, stage by stage
    cyc = d.cycle_counter()
    _, _, _, ru = d.export_philox_noise(cyc)
    d.update_measurement_grid(meas)
    d.update_pose(0.0, 0.4 * 4, 0.0)
    before = d.get_particles()
    d.particle_prediction(0.1)
    pred = d.get_particles()
    assert np.all(pred.grid_cell_idx >= 0) and np.all(pred.grid_cell_idx < C)
    inside = (pred.state[:, 0] >= 0) & (pred.state[:, 0] <= gs - 1) & (pred.state[:, 1] >= 0) & (pred.state[:, 1] <= gs - 1)
    assert np.all(pred.weight[~inside] == 0)
    assert np.array_equal(pred.weight[inside], np.float32(0.99) * before.weight[inside])
    cx = np.clip(pred.state[:, 0].astype(np.int32), 0, gs - 1)  # trunc toward zero like the device
    cy = np.clip(pred.state[:, 1].astype(np.int32), 0, gs - 1)
    assert np.array_equal(pred.grid_cell_idx, cx + gs * cy)

    d.particle_assignment()
    srt = d.get_particles()
    keys = srt.grid_cell_idx
    assert np.all(np.diff(keys) >= 0), "particles are not sorted by cell"
    order = np.argsort(pred.grid_cell_idx, kind="stable")
    assert np.array_equal(srt.state.view(np.uint32), pred.state[order].view(np.uint32)), "the sort is not stable"
    start, end = d.get_cell_ranges()
    occ_cells = np.unique(keys)
    first = np.searchsorted(keys, occ_cells, side="left")
    last = np.searchsorted(keys, occ_cells, side="right") - 1
    assert np.array_equal(start[occ_cells], first) and np.array_equal(end[occ_cells], last)
    empty = np.ones(C, bool)
    empty[occ_cells] = False
    assert np.all(start[empty] == -1)

    d.grid_cell_occupancy_update(0.1)
    g = d.get_grid_cells()
    sums = np.bincount(keys, weights=srt.weight.astype(np.float64), minlength=C)
    assert np.all(np.abs(g["pred_occ_mass"] - np.minimum(sums, 1.0)) <= 2e-7 * np.maximum(np.minimum(sums, 1.0), 1e-30) + 1e-30)
    assert np.all(g["occ_mass"] + g["free_mass"] <= 1.0 + 1e-5)
    assert np.allclose(g["occ_mass"], g["pers_occ_mass"] + g["new_born_occ_mass"], atol=2e-7)
    assert np.array_equal(g["start_idx"], start) and np.array_equal(np.where(start >= 0, g["end_idx"], -1), np.where(start >= 0, end, -1))

    d.update_persistent_particles()
    wa = d.get_weight_array().astype(np.float64)
    per_cell = np.bincount(keys, weights=wa, minlength=C)
    # (atol: a border cell holding only zero-weight out-of-grid particles has rho_p = occ - rho_b = -3e-8, mass_update.cu:84-87)
    assert np.allclose(per_cell[occ_cells], g["pers_occ_mass"][occ_cells], rtol=1e-5, atol=1e-7)

    d.initialize_new_particles()
    bp = d.get_birth_particles()
    born = d.get_born_masses().astype(np.float64)
    per_cell_b = np.bincount(bp.grid_cell_idx, weights=bp.weight.astype(np.float64), minlength=C)
    owners = np.nonzero(per_cell_b > 0)[0]
    assert np.allclose(per_cell_b[owners], born[owners], rtol=1e-5, atol=1e-12)
    assert np.all(np.diff(bp.grid_cell_idx[bp.weight > 0]) >= 0)  # slots are handed out in cell order

    d.resampling()
    cdf = d.get_joint_weight_accum()
    anc = d.get_resampled_indices()
    joint = np.concatenate([wa, bp.weight.astype(np.float64)])
    assert np.allclose(cdf, np.cumsum(joint), rtol=1e-12)
    assert np.all(np.diff(cdf) >= 0)
    total = cdf[-1]
    r = (np.arange(N, dtype=np.float64) + np.float64(ru[0])) * (total / N)
    assert np.array_equal(anc, np.minimum(np.searchsorted(cdf, r, side="left"), N + B - 1)), "ancestors != lower_bound(cdf, offsets)"
    new = d.get_particles()
    pers = anc < N
    assert np.array_equal(new.state[pers].view(np.uint32), srt.state[anc[pers]].view(np.uint32))
    assert np.array_equal(new.state[~pers].view(np.uint32), bp.state[anc[~pers] - N].view(np.uint32))
    assert np.array_equal(new.grid_cell_idx[pers], keys[anc[pers]])
    assert np.all(new.weight == np.float32(total) / np.float32(N))
    # systematic resampling: every ancestor's offspring count is the floor or ceil of its expectation
    counts = np.bincount(anc, minlength=N + B)
    assert np.all(np.abs(counts - joint / total * N) < 1.0 + 1e-6)

    # --- determinism: a second handle with the same seed and inputs ends in the same state, bit for bit
    e = gpu.DOGM(p)
    e.set_options(seed=4242, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
    for c in range(5):
        e.update_grid(meas, 0.0, 0.4 * c, 0.0, 0.1, device=False)
    assert np.array_equal(e.get_particles().block, new.block)
    assert np.array_equal(e.get_grid_cells().view(np.uint8), d.get_grid_cells().view(np.uint8))


def test_full_size_cycle_against_oracle(gpu, orc):
    """One cycle at the headline size against the CPU oracle (exported Philox noise)."""
    rng = np.random.default_rng(52)
    p, po = make_params(gpu, SIZE, RES, N, B), make_params(orc, SIZE, RES, N, B)
    d = gpu.DOGM(p)
    d.set_options(seed=7, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
    o = orc.OracleDOGM(po, resample_mode=orc.RESAMPLE_SYSTEMATIC)
    meas = scene_meas(gpu, rng)
    for c in range(2):
        pn, bn, iv, ru = d.export_philox_noise(d.cycle_counter())
        o.set_noise(pn, bn, iv, ru)
        d.update_grid(meas, 0.0, 0.4 * c, 0.0, 0.1, device=False)
        o.update_grid(meas.view(orc.MEAS_CELL_DTYPE), 0.0, 0.4 * c, 0.0, 0.1)
        q, e = d.get_particles(), o.particles
        assert np.array_equal(q.grid_cell_idx, e.grid_cell_idx)
        assert np.array_equal(q.state.view(np.uint32), e.state.view(np.uint32))
        assert np.array_equal(d.get_resampled_indices(), o.resampled_idx)
        g, ge = d.get_grid_cells(), o.grid_cells
        for f in ("occ_mass", "free_mass", "pers_occ_mass", "new_born_occ_mass"):
            assert np.allclose(g[f], ge[f], rtol=1e-6, atol=1e-9), f
        assert np.allclose(g["mean_x_vel"], ge["mean_x_vel"], rtol=1e-4, atol=3e-3)


def test_no_birth_particles_and_single_particle(gpu, orc):
    """Edge sizes: B = 0 (no birth set at all) and N = 1."""
    from conftest import cycle_noise, synthetic_meas

    for n, b in ((3000, 0), (1, 1)):
        rng = np.random.default_rng(53)
        p, po = make_params(gpu, 8.0, 0.5, n, b), make_params(orc, 8.0, 0.5, n, b)
        d = gpu.DOGM(p)
        d.set_options(noise_mode=gpu.NOISE_INJECTED, resample_mode=gpu.RESAMPLE_INJECTED)
        o = orc.OracleDOGM(po)
        meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, d.grid_size, rng)
        for c in range(3):
            pn, bn, iv, ru = cycle_noise(rng, n, b, p)
            d.set_noise(pn, bn if b else None, iv, ru)
            o.set_noise(pn, bn if b else None, iv, ru)
            d.update_grid(meas, 0.0, 0.6 * c, 0.0, 0.1, device=False)
            o.update_grid(meas.view(orc.MEAS_CELL_DTYPE), 0.0, 0.6 * c, 0.0, 0.1)
            assert np.array_equal(d.get_particles().grid_cell_idx, o.particles.grid_cell_idx)
            assert np.array_equal(d.get_resampled_indices(), o.resampled_idx)
            assert np.allclose(d.get_grid_cells()["occ_mass"], o.grid_cells["occ_mass"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("vx,vy,device_meas", [(0.0, 4.0, False), (3.0, -2.0, True), (0.0, 0.0, True)])
def test_quiet_block_shortcut_is_bit_identical_to_the_dense_path(gpu, monkeypatch, vx, vy, device_meas):
    """The cell kernel skips blocks of 256 cells in which nothing happens for the third cycle in a row (no particles, no
    measurement, no free mass left; most of the grid outside the sensor's field of view).  With the shortcut switched off
    (DOGM_B200_NO_QUIET=1) every cell is computed and stored every cycle: grid cells, measurement copy, born masses, particles
    and birth particles must be the same bits, with ego-motion shifts in any direction and both ways of passing the grid."""
    size, res, n, b = 102.4, 0.2, 300_000, 30_000  # 512 x 512 cells
    rng = np.random.default_rng(5)
    scans = [scene_meas(gpu, rng, 200, size, res) for _ in range(3)]
    out = []
    for off in ("1", "0"):
        monkeypatch.setenv("DOGM_B200_NO_QUIET", off)
        d = gpu.DOGM(make_params(gpu, size, res, n, b))
        d.set_options(seed=99, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
        dev = gpu.device_alloc(scans[0].nbytes)
        snaps = []
        for c in range(9):
            m = scans[c % 3]
            if device_meas:
                gpu.memcpy_h2d(dev, m)
                d.update_grid(dev, vx * 0.1 * (c + 1), vy * 0.1 * (c + 1), 0.0, 0.1, device=True)
            else:
                d.update_grid(m, vx * 0.1 * (c + 1), vy * 0.1 * (c + 1), 0.0, 0.1, device=False)
            if c in (3, 8):
                snaps.append((d.get_grid_cells().copy(), d.get_measurement_cells().copy(), d.get_born_masses().copy(),
                              d.get_particles().block.copy(), d.get_birth_particles().block.copy()))
        gpu.device_free(dev)
        d.close()
        out.append(snaps)
    monkeypatch.delenv("DOGM_B200_NO_QUIET", raising=False)
    for dense, quick in zip(*out):
        for x, y in zip(dense, quick):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    # the test is not vacuous: a large part of the grid sees no measurement at all
    cells = out[0][-1][0]
    assert np.mean((cells["occ_mass"] == 0) & (cells["free_mass"] == 0)) > 0.1
