"""CPU tests of the oracle (oracle/dogm_oracle.c): the reference's two hot-path tests restated, plus
property checks of every stage against independent numpy formulas."""
import numpy as np
import pytest

from conftest import cycle_noise, make_params, synthetic_meas


def spec_params(orc, **over):
    # the Params both dogm_spec.cpp cases use (test/dogm_spec.cpp:12-22, 70-80)
    kw = dict(
        persistence_prob=0.5,
        stddev_process_noise_position=0.0,
        stddev_process_noise_velocity=0.0,
        birth_prob=0.02,
        stddev_velocity=10.0,
        init_max_velocity=30.0,
        freespace_discount=0.01,
    )
    kw.update(over)
    return make_params(orc, 10.0, 1.0, 2, 1, **kw)


def test_spec_predict(orc):
    """DOGM.Predict, test/dogm_spec.cpp:68-103: state' == state + dt*(vx, vy, 0, 0) and weight' == weight * p_S, exactly."""
    o = orc.OracleDOGM(spec_params(orc))
    p = o.particles
    p.state[:] = np.array([[3.25, 4.5, 1.7, -2.3], [7.125, 1.0625, -30.0, 12.5]], np.float32)
    p.weight[:] = np.array([0.37, 0.63], np.float32)
    old_state, old_w = p.state.copy(), p.weight.copy()
    dt = np.float32(0.1)
    o.particle_prediction(float(dt))
    pred = old_state.copy()
    pred[:, 0] = old_state[:, 0] + dt * old_state[:, 2]
    pred[:, 1] = old_state[:, 1] + dt * old_state[:, 3]
    q = o.particles
    assert np.array_equal(q.state, pred)
    assert np.array_equal(q.weight, old_w * np.float32(0.5))
    # cell index = trunc(x) + gs * trunc(y)
    assert list(q.grid_cell_idx) == [int(pred[0, 0]) + 10 * int(pred[0, 1]), int(pred[1, 0]) + 10 * int(pred[1, 1])]


def test_spec_ego_motion_compensation(orc):
    """DOGM.EgoMotionCompensation, test/dogm_spec.cpp:10-66: first pose adopted; sub-cell motion ignored; a +3 m move
    updates the pose and moves particles by +3 cells in x (moveParticlesKernel subtracts x_move = -3)."""
    o = orc.OracleDOGM(spec_params(orc))
    p = o.particles
    p.state[:] = np.array([[3.25, 4.5, 0.0, 0.0], [5.5, 5.5, 0.0, 0.0]], np.float32)
    old = p.state.copy()
    o.update_pose(10.0, 10.0, 0.0)
    assert o.position[:2] == (10.0, 10.0)
    assert np.array_equal(o.particles.state, old)
    o.update_pose(10.5, 10.5, 0.0)
    assert o.position[:2] == (10.0, 10.0)
    assert np.array_equal(o.particles.state, old)
    o.update_pose(13.0, 10.0, 0.0)
    assert o.position[:2] == (13.0, 10.0)
    moved = old.copy()
    moved[:, 0] += 3.0
    assert np.array_equal(o.particles.state, moved)


def test_ego_motion_grid_shift_rule(orc):
    """moveMapKernel, ego_motion_compensation.cu:25-43: new[x,y] = old[x+xm, y+ym] iff 0 < x+xm < gs and 0 < y+ym < gs
    (strict), everything else zero-filled (dogm.cu:185)."""
    o = orc.OracleDOGM(make_params(orc, 8.0, 1.0, 4, 2))
    gs = o.grid_size
    g = o.grid_cells
    g["free_mass"] = np.arange(gs * gs, dtype=np.float32) + 1.0
    old = g["free_mass"].reshape(gs, gs).copy()
    o.update_pose(0.0, 0.0, 0.0)
    o.update_pose(2.0, -3.0, 0.0)  # x_move = -2, y_move = +3
    new = o.grid_cells["free_mass"].reshape(gs, gs)
    exp = np.zeros_like(old)
    for y in range(gs):
        for x in range(gs):
            nx, ny = x - 2, y + 3
            if 0 < nx < gs and 0 < ny < gs:
                exp[y, x] = old[ny, nx]
    assert np.array_equal(new, exp)
    assert np.all(o.grid_cells["start_idx"][exp.ravel() == 0] == 0)  # vacated cells are all-zero, indices included


def test_out_of_grid_particles_are_killed_and_clamped(orc):
    o = orc.OracleDOGM(make_params(orc, 10.0, 1.0, 4, 1, stddev_process_noise_position=0.0, stddev_process_noise_velocity=0.0))
    p = o.particles
    p.state[:] = np.array([[-0.5, 3.0, 0, 0], [9.5, 3.0, 0, 0], [4.0, 12.0, 0, 0], [9.0, 9.0, 0, 0]], np.float32)
    p.weight[:] = 0.25
    o.particle_prediction(0.0)
    q = o.particles
    assert list(q.weight) == [0.0, 0.0, 0.0, np.float32(0.25) * np.float32(0.99)]  # x > gs-1 kills too (predict.cu:41)
    assert list(q.grid_cell_idx) == [0 + 30, 9 + 30, 4 + 90, 9 + 90]


def run_cycles(orc, o, rng, cycles, meas, p, dt=0.1, ego=(0.0, 0.4)):
    for c in range(cycles):
        pn, bn, iv, ru = cycle_noise(rng, o.particle_count, o.new_born_particle_count, p)
        o.set_noise(pn, bn, iv, ru)
        o.update_grid(meas, ego[0] * c, ego[1] * c, 0.0, dt)


def test_assignment_is_a_stable_sort(orc):
    rng = np.random.default_rng(1)
    p = make_params(orc, 12.0, 0.5, 5000, 500)
    o = orc.OracleDOGM(p)
    meas = synthetic_meas(orc.MEAS_CELL_DTYPE, o.grid_size, rng)
    run_cycles(orc, o, rng, 2, meas, p)
    pn, bn, iv, ru = cycle_noise(rng, o.particle_count, o.new_born_particle_count, p)
    o.set_noise(pn, bn, iv, ru)
    o.particle_prediction(0.1)
    before = o.particles.copy()
    o.particle_assignment()
    after = o.particles
    order = np.argsort(before.grid_cell_idx, kind="stable")
    assert np.array_equal(after.grid_cell_idx, before.grid_cell_idx[order])
    assert np.array_equal(after.state, before.state[order])
    assert np.array_equal(after.weight, before.weight[order])
    g = o.grid_cells
    keys = after.grid_cell_idx
    for c in np.unique(keys):
        idx = np.nonzero(keys == c)[0]
        assert g["start_idx"][c] == idx[0] and g["end_idx"][c] == idx[-1]
    empty = np.setdiff1d(np.arange(o.grid_cell_count), np.unique(keys))
    assert np.all(g["start_idx"][empty] == -1) and np.all(g["end_idx"][empty] == -1)


def ds_update_f64(m_occ_pred, free_prev, z_free, z_occ, alpha, p_B):
    """Independent float64 statement of mass_update.cu:16-39."""
    m_free_pred = min(alpha * free_prev, 1.0 - m_occ_pred)
    unk = 1.0 - m_occ_pred - m_free_pred
    zunk = 1.0 - z_free - z_occ
    K = m_free_pred * z_occ + m_occ_pred * z_free
    occ = (m_occ_pred * zunk + unk * z_occ + m_occ_pred * z_occ) / (1.0 - K)
    fre = (m_free_pred * zunk + unk * z_free + m_free_pred * z_free) / (1.0 - K)
    rho_b = occ * p_B * (1.0 - m_occ_pred) / (m_occ_pred + p_B * (1.0 - m_occ_pred))
    return occ, fre, rho_b, occ - rho_b


@pytest.mark.parametrize("sum_mode", [0, 1])
def test_cycle_against_float64_formulas(orc, sum_mode):
    rng = np.random.default_rng(2)
    p = make_params(orc, 16.0, 0.5, 20000, 2000)
    o = orc.OracleDOGM(p, sum_mode=sum_mode)
    gs = o.grid_size
    meas = synthetic_meas(orc.MEAS_CELL_DTYPE, gs, rng)
    run_cycles(orc, o, rng, 3, meas, p)
    # one more cycle, stage by stage
    pn, bn, iv, ru = cycle_noise(rng, o.particle_count, o.new_born_particle_count, p)
    o.set_noise(pn, bn, iv, ru)
    o.update_measurement_grid(meas)
    o.update_pose(0.0, 1.2, 0.0)
    free_prev = o.grid_cells["free_mass"].astype(np.float64).copy()
    o.particle_prediction(0.1)
    o.particle_assignment()
    keys = o.particles.grid_cell_idx.copy()
    w_pred = o.particles.weight.astype(np.float64).copy()
    o.grid_cell_occupancy_update(0.1)
    g = o.grid_cells.copy()
    sums = np.bincount(keys, weights=w_pred, minlength=o.grid_cell_count)
    alpha = float(np.float32(0.01) ** np.float32(0.1))
    tol = 2e-4 if sum_mode == 1 else 2e-5
    for c in rng.choice(o.grid_cell_count, size=300, replace=False):
        m = min(sums[c], 1.0)
        occ, fre, rho_b, rho_p = ds_update_f64(m, free_prev[c], float(meas["free_mass"][c]), float(meas["occ_mass"][c]), alpha, 0.02)
        assert g["occ_mass"][c] == pytest.approx(occ, rel=tol, abs=tol)
        assert g["free_mass"][c] == pytest.approx(fre, rel=tol, abs=tol)
        assert g["new_born_occ_mass"][c] == pytest.approx(rho_b, rel=tol, abs=tol)
        assert g["pers_occ_mass"][c] == pytest.approx(rho_p, rel=tol, abs=tol)
    assert np.allclose(g["occ_mass"], g["pers_occ_mass"] + g["new_born_occ_mass"], atol=1e-6)
    assert np.all(g["occ_mass"] + g["free_mass"] <= 1.0 + 1e-5)

    o.update_persistent_particles()
    wa = o.weight_array.astype(np.float64)
    per_cell = np.bincount(keys, weights=wa, minlength=o.grid_cell_count)
    occ_cells = np.unique(keys)
    # the updated persistent weights of a cell add up to its persistent occupancy mass (Nuss et al., eq. 47)
    # (the reference's float scan-difference sums, sum_mode 1, are only good to ~1e-3 relative already at N = 2e4)
    assert np.allclose(
        per_cell[occ_cells], g["pers_occ_mass"][occ_cells].astype(np.float64), rtol=1e-2 if sum_mode else 1e-6, atol=1e-6
    )

    o.initialize_new_particles()
    bp = o.birth_particles
    per_cell_b = np.bincount(bp.grid_cell_idx, weights=bp.weight.astype(np.float64), minlength=o.grid_cell_count)
    owners = np.nonzero(per_cell_b > 0)[0]
    assert np.allclose(per_cell_b[owners], g["new_born_occ_mass"][owners], rtol=1e-4, atol=1e-7)
    # birth position: x = col + .5, y = row + col/gs + .5 (float division quirk, init_new_particles.cu:173)
    j = bp.grid_cell_idx
    assert np.array_equal(bp.state[:, 0], (j % gs).astype(np.float32) + np.float32(0.5))
    assert np.array_equal(bp.state[:, 1], j.astype(np.float32) / np.float32(gs) + np.float32(0.5))
    assert np.array_equal(bp.state[:, 2:], bn)

    o.statistical_moments()
    g2 = o.grid_cells
    vx, vy = o.particles.state[:, 2].astype(np.float64), o.particles.state[:, 3].astype(np.float64)
    for c in occ_cells[:: max(1, len(occ_cells) // 100)]:
        sel = keys == c
        rho_p = float(g2["pers_occ_mass"][c])
        if rho_p <= 0:
            continue
        mx = (wa[sel] * vx[sel]).sum() / rho_p
        vxx = (wa[sel] * vx[sel] ** 2).sum() / rho_p - mx * mx
        assert g2["mean_x_vel"][c] == pytest.approx(mx, rel=1e-3, abs=5e-2 if sum_mode else 1e-3)
        assert g2["var_x_vel"][c] == pytest.approx(vxx, rel=2e-3, abs=2.0 if sum_mode else 5e-2)

    o.resampling()
    n = o.particle_count
    joint = np.concatenate([o.weight_array, o.birth_particles.weight]).astype(np.float64)
    cdf = np.cumsum(joint)
    assert np.allclose(o.joint_weight_accum, cdf, rtol=1e-4 if sum_mode else 1e-12)
    if sum_mode == 0:
        draws = (np.float32(o.joint_max) * ru).astype(np.float64)
        exp = np.minimum(np.searchsorted(o.joint_weight_accum, draws, side="left"), n + o.new_born_particle_count - 1)
        assert np.array_equal(o.resampled_idx, exp)
    anc = o.resampled_idx
    assert np.all(np.diff(anc) >= 0)
    nxt = o.particles_next
    pers = anc < n
    assert np.array_equal(nxt.state[pers], o.particles.state[anc[pers]])
    assert np.array_equal(nxt.state[~pers], o.birth_particles.state[anc[~pers] - n])
    assert np.all(nxt.weight == np.float32(o.joint_max) / np.float32(n))


def test_resample_modes_systematic_and_stratified(orc):
    rng = np.random.default_rng(3)
    p = make_params(orc, 8.0, 0.5, 3000, 300)
    for mode in (orc.RESAMPLE_SYSTEMATIC, orc.RESAMPLE_STRATIFIED):
        o = orc.OracleDOGM(p, resample_mode=mode)
        meas = synthetic_meas(orc.MEAS_CELL_DTYPE, o.grid_size, rng)
        pn, bn, iv, ru = cycle_noise(rng, o.particle_count, o.new_born_particle_count, p)
        u = rng.uniform(0, 1, size=o.particle_count).astype(np.float32)
        o.set_noise(pn, bn, iv, u)
        o.update_grid(meas, 0.0, 0.0, 0.0, 0.1)
        cdf = o.joint_weight_accum
        total = cdf[-1]
        i = np.arange(o.particle_count, dtype=np.float64)
        frac = u[0] if mode == orc.RESAMPLE_SYSTEMATIC else u.astype(np.float64)
        r = (i + frac) * (total / o.particle_count)
        exp = np.minimum(np.searchsorted(cdf, r, side="left"), cdf.size - 1)
        assert np.array_equal(o.resampled_idx, exp)
        # systematic resampling: offspring count of every ancestor is floor or ceil of its expected count
        if mode == orc.RESAMPLE_SYSTEMATIC:
            joint = np.diff(np.concatenate([[0.0], cdf]))
            expected = joint / total * o.particle_count
            counts = np.bincount(o.resampled_idx, minlength=cdf.size)
            assert np.all(np.abs(counts - expected) < 1.0 + 1e-6)


def test_first_cycle_initialisation(orc):
    rng = np.random.default_rng(4)
    p = make_params(orc, 10.0, 0.5, 4000, 400)
    o = orc.OracleDOGM(p)
    gs = o.grid_size
    meas = synthetic_meas(orc.MEAS_CELL_DTYPE, gs, rng)
    pn, bn, iv, ru = cycle_noise(rng, o.particle_count, o.new_born_particle_count, p)
    o.set_noise(pn, bn, iv, ru)
    o.update_measurement_grid(meas)
    q = o.particles
    # particles are distributed over cells in proportion to the measured occupancy mass, at cell centres
    counts = np.bincount(q.grid_cell_idx, minlength=gs * gs).astype(np.float64)
    expect = meas["occ_mass"].astype(np.float64) / meas["occ_mass"].sum() * o.particle_count
    assert np.all(np.abs(counts - expect) <= 1.0 + 1e-6)
    assert np.array_equal(q.state[:, 0], (q.grid_cell_idx % gs).astype(np.float32) + np.float32(0.5))
    assert np.array_equal(q.state[:, 1], (q.grid_cell_idx // gs).astype(np.float32) + np.float32(0.5))
    assert np.array_equal(q.state[:, 2:], iv)
    assert np.all(q.weight == np.float32(1.0) / np.float32(o.particle_count))


def test_null_measurement_and_zero_mass(orc):
    """updateGrid(nullptr, ...) as in dogm_spec.cpp:32: the initial measurement grid (0, 0, 1, 1) stays in place and the
    cycle still runs (no occupied mass -> no slots: particles stay in cell 0)."""
    o = orc.OracleDOGM(spec_params(orc))
    o.set_noise(np.zeros((2, 4), np.float32), np.zeros((1, 2), np.float32), np.zeros((2, 2), np.float32), np.array([0.25, 0.75], np.float32))
    o.update_grid(None, 10.0, 10.0, 0.0, 0.0)
    assert o.position[:2] == (10.0, 10.0)
    assert np.all(o.meas_cells["likelihood"] == 1.0) and np.all(o.meas_cells["occ_mass"] == 0.0)
    assert np.all(np.isfinite(o.grid_cells["occ_mass"]))
    assert np.array_equal(o.particles.state[:, :2], np.full((2, 2), 0.5, np.float32))


def test_over_unit_cell_is_renormalised(orc):
    """normalize_weights, mass_update.cu:51-59,79-83."""
    p = make_params(orc, 4.0, 1.0, 8, 2, stddev_process_noise_position=0.0, stddev_process_noise_velocity=0.0, persistence_prob=1.0)
    o = orc.OracleDOGM(p)
    q = o.particles
    q.state[:] = np.array([[1.5, 1.5, 0, 0]] * 6 + [[2.5, 2.5, 0, 0]] * 2, np.float32)
    q.weight[:] = np.array([0.3] * 6 + [0.1] * 2, np.float32)
    o.particle_prediction(0.0)
    o.particle_assignment()
    o.grid_cell_occupancy_update(0.1)
    g = o.grid_cells
    assert g["pred_occ_mass"][1 + 4 * 1] == 1.0
    assert g["pred_occ_mass"][2 + 4 * 2] == pytest.approx(0.2, rel=1e-6)
    assert np.allclose(o.weight_array[:6], 0.3 / 1.8, rtol=1e-6)
    o.update_persistent_particles()
    assert o.weight_array[:6].sum() == pytest.approx(float(g["pers_occ_mass"][5]), rel=1e-5)


def test_search_ancestors_f32_matches_searchsorted(orc):
    rng = np.random.default_rng(5)
    w = rng.uniform(0, 1, 5000).astype(np.float32)
    w[rng.integers(0, 5000, 500)] = 0.0  # zero-weight entries (out-of-grid particles) are never selected
    cdf = np.cumsum(w, dtype=np.float32)
    draws = np.sort(rng.uniform(0, cdf[-1], 4000).astype(np.float32))
    got = orc.search_ancestors_f32(cdf, draws)
    exp = np.minimum(np.searchsorted(cdf, draws, side="left"), cdf.size - 1)
    assert np.array_equal(got, exp)
    assert np.all(w[got] > 0)
