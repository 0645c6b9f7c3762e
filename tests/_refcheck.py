"""Stage-wise comparison of an implementation against snapshots of the reference's own CUDA run
(oracle/ref.py: RefDOGM.run_cycle_verbose), used with live snapshots on the GPU box and with the committed golden
fixtures (tests/golden/ref_*.npz) on the CPU.

What can and cannot be equal: prediction, cell indices and the stable sort are bit-exact; everything
downstream of a thrust::inclusive_scan in the reference (per-cell sums as float prefix differences, birth slot
boundaries, the joint CDF) carries the reference's own float error of a few ulp of the RUNNING TOTAL, so masses are
compared with rtol 1e-4 plus that noise floor, and ancestor indices are compared bit-exactly with the reference's own
CDF and sorted draws fed to the search."""
import numpy as np

EPS32 = float(np.finfo(np.float32).eps)


def split_block(block, n):
    b = np.ascontiguousarray(block).view(np.uint8)
    state = b[: 16 * n].view("<f4").reshape(n, 4)
    idx = b[16 * n : 20 * n].view("<i4")
    weight = b[20 * n : 24 * n].view("<f4")
    assoc = b[24 * n : 25 * n]
    return state, idx, weight, assoc


class OracleAdapter:
    """Drives oracle.OracleDOGM (CPU)."""

    def __init__(self, orc, params):
        self.mod = orc
        self.o = orc.OracleDOGM(params, resample_mode=orc.RESAMPLE_INJECTED)
        self.N, self.B = self.o.particle_count, self.o.new_born_particle_count

    def set_state(self, P0, G0, meas, pose0, have_pose=True):
        st, idx, w, a = split_block(P0, self.N)
        p = self.o.particles
        p.state[:], p.grid_cell_idx[:], p.weight[:], p.associated[:] = st, idx, w, a
        self.o.grid_cells[:] = G0.view(self.mod.GRID_CELL_DTYPE)
        self.o.set_first_measurement_received(True)
        self.o.update_measurement_grid(meas.view(self.mod.MEAS_CELL_DTYPE))
        if have_pose:
            self.o.set_pose(float(pose0[0]), float(pose0[1]), float(pose0[2]))

    def fresh_init(self, meas, iv):
        self.o.set_noise(init_velocity=iv)
        self.o.update_measurement_grid(meas.view(self.mod.MEAS_CELL_DTYPE))
        p = self.o.particles
        return p.state.copy(), p.grid_cell_idx.copy(), p.weight.copy()

    def set_noise(self, pn, bn, iv, ru):
        self.o.set_noise(pn, bn, iv, ru)

    def update_pose(self, x, y, yaw):
        self.o.update_pose(x, y, yaw)

    def predict(self, dt):
        self.o.particle_prediction(dt)
        p = self.o.particles
        return p.state.copy(), p.grid_cell_idx.copy(), p.weight.copy(), p.associated.copy()

    def assignment(self):
        self.o.particle_assignment()
        p, g = self.o.particles, self.o.grid_cells
        return (p.state.copy(), p.grid_cell_idx.copy(), p.weight.copy(), p.associated.copy()), g["start_idx"].copy(), g["end_idx"].copy()

    def occupancy(self, dt):
        self.o.grid_cell_occupancy_update(dt)
        return self.o.grid_cells.copy(), self.o.born_masses.copy()

    def persistent(self):
        self.o.update_persistent_particles()
        return self.o.weight_array.copy(), self.o.grid_cells.copy()

    def birth(self):
        self.o.initialize_new_particles()
        b = self.o.birth_particles
        return (b.state.copy(), b.grid_cell_idx.copy(), b.weight.copy(), b.associated.copy()), self.o.grid_cells.copy()

    def moments(self):
        self.o.statistical_moments()
        return self.o.grid_cells.copy()

    def resampling(self):
        self.o.resampling()
        n = self.o.particles_next
        return self.o.joint_weight_accum.copy(), self.o.resampled_idx.copy(), (n.state.copy(), n.grid_cell_idx.copy(), n.weight.copy(), n.associated.copy())

    def search_f32(self, cdf, draws):
        return self.mod.search_ancestors_f32(cdf, draws)

    def position(self):
        return self.o.position


class GpuAdapter:
    """Drives dogm_b200.DOGM (the CUDA library through its C ABI)."""

    def __init__(self, gpu, params):
        self.mod = gpu
        self.params = params
        self.d = gpu.DOGM(params)
        self.d.set_options(noise_mode=gpu.NOISE_INJECTED, resample_mode=gpu.RESAMPLE_INJECTED)
        self.N, self.B = self.d.particle_count, self.d.new_born_particle_count

    def set_state(self, P0, G0, meas, pose0, have_pose=True):
        self.d.set_particles(self.mod.ParticlesSoA(self.N, np.ascontiguousarray(P0).view(np.uint8).copy()))
        self.d.set_grid_cells(np.ascontiguousarray(G0).view(self.mod.GRID_CELL_DTYPE))
        self.d.set_measurement_cells(np.ascontiguousarray(meas).view(self.mod.MEAS_CELL_DTYPE))
        if have_pose:
            self.d.set_pose(float(pose0[0]), float(pose0[1]), float(pose0[2]))

    def fresh_init(self, meas, iv):
        d = self.mod.DOGM(self.params)
        d.set_options(noise_mode=self.mod.NOISE_INJECTED, resample_mode=self.mod.RESAMPLE_INJECTED)
        d.set_noise(init_velocity=iv)
        d.update_measurement_grid(np.ascontiguousarray(meas).view(self.mod.MEAS_CELL_DTYPE))
        p = d.get_particles()
        out = p.state.copy(), p.grid_cell_idx.copy(), p.weight.copy()
        d.close()
        return out

    def set_noise(self, pn, bn, iv, ru):
        self.d.set_noise(pn, bn, iv, ru)

    def update_pose(self, x, y, yaw):
        self.d.update_pose(x, y, yaw)

    def _parts(self, p):
        return p.state.copy(), p.grid_cell_idx.copy(), p.weight.copy(), p.associated.copy()

    def predict(self, dt):
        self.d.particle_prediction(dt)
        return self._parts(self.d.get_particles())

    def assignment(self):
        self.d.particle_assignment()
        s, e = self.d.get_cell_ranges()
        e = np.where(s >= 0, e, -1)
        return self._parts(self.d.get_particles()), s, e

    def occupancy(self, dt):
        self.d.grid_cell_occupancy_update(dt)
        return self.d.get_grid_cells(), self.d.get_born_masses()

    def persistent(self):
        self.d.update_persistent_particles()
        return self.d.get_weight_array(), self.d.get_grid_cells()

    def birth(self):
        self.d.initialize_new_particles()
        return self._parts(self.d.get_birth_particles()), self.d.get_grid_cells()

    def moments(self):
        self.d.statistical_moments()
        return self.d.get_grid_cells()

    def resampling(self):
        self.d.resampling()
        return self.d.get_joint_weight_accum(), self.d.get_resampled_indices(), self._parts(self.d.get_particles())

    def search_f32(self, cdf, draws):
        return self.d.search_ancestors_f32(cdf, draws)

    def position(self):
        return self.d.get_position_x(), self.d.get_position_y(), self.d.get_yaw()


def _close(a, b, rtol, atol):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) <= rtol * np.abs(b.astype(np.float64)) + atol


def check_first_cycle_init(impl, snap, meas, N, gs):
    """initializeParticles (dogm.cu:217-241): particles per cell proportional to the measured occupancy (slot borders
    may move by one against the reference's float scan), positions at cell centres, velocities = the injected draws."""
    st, idx, w = impl.fresh_init(meas, snap["iv"])
    rst, ridx, rw, _ = split_block(snap["P0"], N)
    assert np.array_equal(st[:, 2:], rst[:, 2:]), "first-cycle velocities differ"
    assert np.array_equal(w, rw), "first-cycle weights differ"
    mine = np.bincount(idx, minlength=gs * gs).astype(np.int64)
    ref = np.bincount(ridx, minlength=gs * gs).astype(np.int64)
    assert np.max(np.abs(np.cumsum(mine) - np.cumsum(ref))) <= 1, "first-cycle slot borders differ by more than one slot"
    assert np.array_equal(st[:, 0], (idx % gs).astype(np.float32) + np.float32(0.5))
    assert np.array_equal(st[:, 1], (idx // gs).astype(np.float32) + np.float32(0.5))
    return float(np.mean(idx == ridx))


def f64_truth(snap, meas, ridx, rw, rst, dt, p_B, freespace_discount):
    """The cycle's per-cell quantities in float64 from the reference's own inputs of the stage (its sorted predicted
    particles, its shifted grid, the measurement grid): the arbiter between this implementation and the reference, whose
    per-cell sums are differences of a float32 running total.  mass_update.cu:61-93,
    update_persistent_particles.cu:49-87, statistical_moments.cu:78-106."""
    C = snap["G2"].size
    w = rw.astype(np.float64)
    m_sum = np.bincount(ridx, weights=w, minlength=C)
    m_occ = np.minimum(m_sum, 1.0)
    alpha = float(np.power(np.float32(freespace_discount), np.float32(dt)))
    free_prev = snap["G2"]["free_mass"].astype(np.float64)  # after the reference's updatePose has moved the map
    z_free, z_occ = meas["free_mass"].astype(np.float64), meas["occ_mass"].astype(np.float64)
    m_free = np.minimum(alpha * free_prev, 1.0 - m_occ)
    unknown = 1.0 - m_occ - m_free
    meas_unknown = 1.0 - z_free - z_occ
    K = m_free * z_occ + m_occ * z_free
    occ_up = (m_occ * meas_unknown + unknown * z_occ + m_occ * z_occ) / (1.0 - K)
    free_up = (m_free * meas_unknown + unknown * z_free + m_free * z_free) / (1.0 - K)
    rho_b = occ_up * p_B * (1.0 - m_occ) / (m_occ + p_B * (1.0 - m_occ))
    rho_p = occ_up - rho_b
    with np.errstate(divide="ignore", invalid="ignore"):
        share = np.where(m_sum[ridx] > 0, w / m_sum[ridx], 0.0)  # a particle's share of its cell (over-unit cells are normalised)
        weights = rho_p[ridx] * share
        inv = np.where(m_sum > 0, 1.0 / m_sum, 0.0)
    vx, vy = rst[:, 2].astype(np.float64), rst[:, 3].astype(np.float64)
    mean_x = np.bincount(ridx, weights=w * vx, minlength=C) * inv
    mean_y = np.bincount(ridx, weights=w * vy, minlength=C) * inv
    speed = np.bincount(ridx, weights=w * np.hypot(vx, vy), minlength=C) * inv
    return {"pred_occ_mass": m_occ, "occ_mass": occ_up, "free_mass": free_up, "new_born_occ_mass": rho_b, "pers_occ_mass": rho_p,
            "weights": weights,
            "mean_x_vel": mean_x, "mean_y_vel": mean_y, "speed": speed, "m_sum": m_sum}


def check_cycle(impl, snap, meas, x, y, yaw, dt, first_cycle=False, p_B=0.02, freespace_discount=0.01):
    """Runs one cycle of `impl` from the reference's starting state and compares stage by stage.  Returns stats.
    first_cycle: the reference has not received a pose yet (its yaw member is even uninitialised, dogm.cu:33-37)."""
    N, B = impl.N, impl.B
    stats = {}
    impl.set_state(snap["P0"], snap["G0"], meas, snap["pose0"], have_pose=not first_cycle)
    impl.set_noise(snap["pn"], snap["bn"], snap["iv"], snap["ru"])

    # --- pose + prediction: bit-exact (predict.cu:16-52, ego_motion_compensation.cu:16-23)
    impl.update_pose(x, y, yaw)
    st, idx, w, _ = impl.predict(dt)
    rst, ridx, rw, _ = split_block(snap["P1"], N)
    assert np.array_equal(idx, ridx), "particle-to-cell indices differ from the reference"
    assert np.array_equal(st.view(np.uint32), rst.view(np.uint32)), "predicted states differ from the reference"
    assert np.array_equal(w.view(np.uint32), rw.view(np.uint32)), "predicted weights differ from the reference"
    assert impl.position() == tuple(float(v) for v in snap["pose7"]), "pose bookkeeping differs"

    # --- assignment: the stable sort permutation and the cell ranges are bit-exact (dogm.cu:262-281)
    (st, idx, w, _), cs, ce = impl.assignment()
    rst, ridx, rw, _ = split_block(snap["P2"], N)
    assert np.array_equal(idx, ridx), "sorted cell indices differ"
    assert np.array_equal(st.view(np.uint32), rst.view(np.uint32)), "sorted states differ (sort not stable?)"
    assert np.array_equal(w.view(np.uint32), rw.view(np.uint32)), "sorted weights differ"
    assert np.array_equal(cs, snap["G2"]["start_idx"]) and np.array_equal(ce, snap["G2"]["end_idx"]), "cell ranges differ"

    # --- occupancy update: masses within 1e-4 relative + the reference's scan noise floor
    total_w = float(np.sum(rw.astype(np.float64)))
    floor = 8.0 * EPS32 * max(total_w, 1.0)
    g, born = impl.occupancy(dt)
    G3 = snap["G3"]
    worst = 0.0
    # d(occ, free)/d(pred) is O(1); the born / persistent split divides by (pred + p_B (1 - pred)) >= p_B
    amp = {"pred_occ_mass": 1.0, "occ_mass": 4.0, "free_mass": 4.0, "new_born_occ_mass": 4.0 / p_B, "pers_occ_mass": 4.0 / p_B}
    for f in ("pred_occ_mass", "occ_mass", "free_mass", "new_born_occ_mass", "pers_occ_mass"):
        ok = _close(g[f], G3[f], 1e-4, floor * amp[f])
        assert np.all(ok), f"{f}: {np.count_nonzero(~ok)} cells outside tolerance, worst {np.max(np.abs(g[f] - G3[f]))}"
        denom = np.maximum(np.abs(G3[f].astype(np.float64)), 1e-3)
        worst = max(worst, float(np.max(np.abs(g[f].astype(np.float64) - G3[f]) / denom)))
    stats["mass_worst_rel"] = worst
    assert np.all(_close(born, snap["born3"], 1e-4, floor * 4.0 / p_B))
    # arbitration: the predicted occupancy of a cell is the float64 sum of its particles' weights,
    # capped at 1; this implementation must be within float rounding of it, whatever the reference's scan error is
    truth = np.minimum(np.bincount(ridx, weights=rw.astype(np.float64), minlength=g.size), 1.0)
    err_mine = np.abs(g["pred_occ_mass"].astype(np.float64) - truth)
    err_ref = np.abs(G3["pred_occ_mass"].astype(np.float64) - truth)
    assert np.all(err_mine <= 2.0 * EPS32 * np.maximum(truth, 1e-30)), "predicted occupancy is not the rounded exact cell sum"
    stats["pred_occ_err_vs_f64"] = {"mine_max": float(err_mine.max()), "ref_max": float(err_ref.max())}
    # the same arbitration for every mass the stage produces: this implementation within 1e-4 relative of the float64 truth
    # (the absolute term covers the cancellation in pers = occ - born of nearly empty cells); the reference's distance from
    # the truth is only reported
    t64 = f64_truth(snap, meas, ridx, rw, rst, dt, p_B, freespace_discount)
    arb = {}
    for f in ("occ_mass", "free_mass", "new_born_occ_mass", "pers_occ_mass"):
        mine = np.abs(g[f].astype(np.float64) - t64[f])
        refe = np.abs(G3[f].astype(np.float64) - t64[f])
        bad = mine > 1e-4 * np.abs(t64[f]) + 1e-6
        assert not np.any(bad), f"{f}: {np.count_nonzero(bad)} cells off the float64 truth, worst {mine.max()}"
        arb[f] = (float(mine.max()), float(refe.max()))
    mine = np.abs(born.astype(np.float64) - t64["new_born_occ_mass"])
    assert np.all(mine <= 1e-4 * t64["new_born_occ_mass"] + 1e-6), "born masses off the float64 truth"
    arb["born_masses"] = (float(mine.max()), float(np.abs(snap["born3"].astype(np.float64) - t64["new_born_occ_mass"]).max()))
    stats["max_abs_err_vs_f64(mine, ref)"] = arb

    # --- persistent weights (update_persistent_particles.cu:49-87)
    wa, g4 = impl.persistent()
    W4 = snap["W4"]
    rel = np.abs(wa.astype(np.float64) - W4) / np.maximum(np.abs(W4.astype(np.float64)), 1e-30)
    per_cell_floor = floor / np.maximum(snap["G3"]["pred_occ_mass"][ridx].astype(np.float64), 1e-12)  # relative error of the cell sum
    ok = rel <= 1e-4 + 6.0 * per_cell_floor
    assert np.all(ok | (W4 == 0)), f"weight_array: {np.count_nonzero(~(ok | (W4 == 0)))} particles outside tolerance, worst {rel[~ok].max() if np.any(~ok) else 0}"
    stats["weight_median_rel"] = float(np.median(rel[W4 > 0])) if np.any(W4 > 0) else 0.0
    # float64 arbitration of the persistent weights: weight = pers_occ_mass of the cell * the particle's share of the cell
    tw = t64["weights"]
    werr = np.abs(wa.astype(np.float64) - tw)
    wtol = 1e-4 * np.abs(tw) + 1e-6 * np.where(t64["m_sum"][ridx] > 0, rw.astype(np.float64) / np.maximum(t64["m_sum"][ridx], 1e-300), 0.0)
    assert np.all(werr <= wtol), f"weight_array: {np.count_nonzero(werr > wtol)} particles off the float64 truth"
    with np.errstate(divide="ignore", invalid="ignore"):
        stats["weight_rel_err_vs_f64(mine, ref)"] = (float(np.nanmax(np.where(tw > 0, werr / tw, 0.0))),
                                                     float(np.nanmax(np.where(tw > 0, np.abs(W4.astype(np.float64) - tw) / tw, 0.0))))

    # --- birth: total born mass per cell is conserved; slot ownership is racy in the reference
    (bst, bidx, bw, bas), g5 = impl.birth()
    rbst, rbidx, rbw, rbas = split_block(snap["BP5"], B)
    assert np.array_equal(bst[:, 2:], rbst[:, 2:]), "birth velocities differ"
    C = g5.size
    mine = np.bincount(bidx, weights=bw.astype(np.float64), minlength=C)
    owners = np.nonzero(mine > 0)[0]  # (a slot no cell owns keeps a stale index with weight 0: not an owner)
    assert np.all(_close(mine[owners], born[owners], 2e-4, 1e-7)), "birth weights of a cell do not add up to its born mass"
    stats["birth_mass_vs_ref"] = float(mine.sum() / max(float(np.sum(rbw.astype(np.float64))), 1e-30))
    stats["birth_slot_match"] = float(np.mean(bidx == rbidx))
    cnt_m = np.cumsum(np.bincount(bidx, minlength=C))
    cnt_r = np.cumsum(np.bincount(rbidx, minlength=C))
    stats["birth_border_max_shift"] = int(np.max(np.abs(cnt_m - cnt_r)))

    # --- moments: mean velocities within 1e-4 relative (+ floor), statistical_moments.cu:78-106
    g6 = impl.moments()
    G6 = snap["G6"]
    occ = (G6["start_idx"] >= 0) & (G6["pers_occ_mass"] > 1e-4)
    # the reference's w*v prefix sums run up to sum |w v| before a cell's segment is taken as a difference
    for f, col in (("mean_x_vel", 2), ("mean_y_vel", 3)):
        run = float(np.sum(np.abs(W4.astype(np.float64) * rst[:, col].astype(np.float64))))
        vel_floor = 8.0 * EPS32 * max(run, 1.0) / np.maximum(G6["pers_occ_mass"][occ].astype(np.float64), 1e-6)
        a, b = g6[f][occ].astype(np.float64), G6[f][occ].astype(np.float64)
        # ... and its weights are normalised with a cell sum from a second, differently rounded scan, so they carry
        # the relative error of that sum (floor / predicted mass of the cell)
        rel_sum = 6.0 * floor / np.maximum(G6["pred_occ_mass"][occ].astype(np.float64), 1e-12)
        ok = np.abs(a - b) <= (1e-4 + rel_sum) * np.abs(b) + vel_floor + 1e-4
        assert ok.size == 0 or np.all(ok), f"{f}: {np.count_nonzero(~ok)} of {ok.size} cells outside tolerance"
    empty = G6["start_idx"] < 0
    assert np.all(g6["mean_x_vel"][empty] == 0) and np.all(g6["var_x_vel"][empty] == 0)
    # float64 arbitration of the mean velocities (cells with a persistent mass): within 1e-4 of the truth, relative to the
    # cell's mean speed (a mean near zero is the difference of large terms)
    sel = (t64["m_sum"] > 0) & (g6["pers_occ_mass"] > 0)
    varb = {}
    for f in ("mean_x_vel", "mean_y_vel"):
        tol = 1e-4 * (np.abs(t64[f][sel]) + t64["speed"][sel]) + 1e-6
        mine_e = np.abs(g6[f][sel].astype(np.float64) - t64[f][sel])
        assert np.all(mine_e <= tol), f"{f}: {np.count_nonzero(mine_e > tol)} cells off the float64 truth, worst {mine_e.max() if mine_e.size else 0}"
        sel_r = sel & (G6["pers_occ_mass"] > 0)
        varb[f] = (float(mine_e.max()) if mine_e.size else 0.0,
                   float(np.abs(G6[f][sel_r].astype(np.float64) - t64[f][sel_r]).max()) if np.any(sel_r) else 0.0)
    stats["mean_vel_abs_err_vs_f64(mine, ref)"] = varb

    # --- resampling: CDF within float tolerance; ancestors bit-exact on the reference's own CDF and draws
    cdf, anc, (nst, nidx, nw, nas) = impl.resampling()
    rcdf = snap["cdf7"].astype(np.float64)
    # persistent part of the CDF; the birth part [N, N+B) inherits the reference's racy slot ownership (it hands a few
    # percent of the slots to empty cells with weight 0, losing born mass) and is only reported
    # every weight may be off by the relative error of its cell sum (see above) and those errors add up along the CDF
    cum_tol = np.cumsum(np.abs(W4.astype(np.float64)) * (1e-4 + 6.0 * per_cell_floor)) + 16.0 * EPS32 * rcdf[:N] + floor * 4
    assert np.all(np.abs(cdf[:N] - rcdf[:N]) <= cum_tol), "joint CDF (persistent part) differs from the reference"
    stats["cdf_birth_part_max_rel"] = float(np.max(np.abs(cdf[N:] - rcdf[N:]) / np.maximum(rcdf[N:], 1e-30))) if B > 0 else 0.0
    got = impl.search_f32(snap["cdf7"], snap["rand7"])
    assert np.array_equal(got, snap["idx7"]), "ancestor indices differ from the reference (same CDF, same draws)"
    # k_resample itself, pinned independently of the C oracle: on the implementation's OWN double CDF and the injected draws
    # the ancestors must be exactly numpy's lower_bound of  float(total) * u  (resampling.cu:45 + clamp).  How often they equal
    # the reference's is only reported: its float CDF and racy birth slots (other born weights, another total) move the
    # offsets, so that agreement says nothing about the search (measured 0.09 - 0.99).
    stats["ancestor_match_own_cdf"] = float(np.mean(anc == snap["idx7"]))
    own_total = np.float32(cdf[-1])
    own_r = (own_total * np.asarray(snap["ru"], np.float32)).astype(np.float64)
    expect = np.minimum(np.searchsorted(cdf, own_r, side="left"), len(cdf) - 1)
    assert np.array_equal(anc, expect), "ancestors are not the lower bounds of the draws on the implementation's own CDF"
    jm = np.float32(snap["joint_max7"][0])
    # the weight total differs by the born mass of the few cells whose slot count differs (racy slot ownership)
    stats["joint_max_vs_ref"] = float(nw[0]) * N / float(jm) if float(jm) > 0 else float("nan")
    stats["joint_max_ref"] = float(jm)
    # the gather itself: next[i] = particle[a_i] or birth[a_i - N] (resampling.cu:49-68), checked on own ancestors
    pers = anc < N
    assert np.array_equal(nst[pers], st[anc[pers]])
    assert np.array_equal(nst[~pers], bst[anc[~pers] - N])
    assert np.array_equal(nidx[pers], idx[anc[pers]])
    return stats
