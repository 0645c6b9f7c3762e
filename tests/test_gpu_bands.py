"""Band mode (include/dogm_b200.h): one grid over several handles, each owning a band of rows, driven in phases by
BandedDOGM (all bands on one GPU here; the exchange steps are the ones a multi-GPU run does over NVLink).
  * one band covering the whole grid reproduces dogm_update_grid bit for bit (same phases, same kernels);
  * with several bands: every particle lies in its band, the global particle count and the total weight are what the
    single-grid run has, the per-band counts follow the mass, and the maps agree statistically with the single-grid run
    (the bands draw other random numbers, so not bit for bit)."""
import numpy as np
import pytest

from conftest import make_params

pytestmark = pytest.mark.gpu

SIZE, RES, N, B = 60.0, 0.2, 400_000, 40_000  # 300 x 300 cells


def scans(gpu, steps, seed=3):
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(SIZE, RES, 120.0, 0.5), SIZE, RES)
    rng = np.random.default_rng(seed)
    out = []
    z = np.full(120, np.inf, np.float32)
    hit = rng.uniform(size=120) < 0.6
    z[hit] = rng.uniform(6.0, 55.0, size=int(hit.sum())).astype(np.float32)
    for s in range(steps):
        z = z.copy()
        z[hit] = np.clip(z[hit] + rng.normal(0, 0.3, int(hit.sum())), 3.0, 58.0).astype(np.float32)
        out.append(gen.generate_grid_host(z))
    gen.close()
    return out


def run_plain(gpu, grids, poses, seed=77):
    d = gpu.DOGM(make_params(gpu, SIZE, RES, N, B))
    d.set_options(seed=seed, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
    for g, (x, y) in zip(grids, poses):
        d.update_grid(g, x, y, 0.0, 0.1, device=False)
    return d


def run_banded(gpu, grids, poses, bands, seed=77, native=True, device_paced=None, devices=None):
    bd = gpu.BandedDOGM(make_params(gpu, SIZE, RES, N, B), bands, seed=seed, native=native, device_paced=device_paced, devices=devices)
    dev = gpu.device_alloc(grids[0].nbytes)
    history = []
    per_dev = {}
    if devices is not None:  # every band reads its rows of the measurement grid from its own GPU
        for d in sorted(set(devices)):
            gpu.set_device(d)
            per_dev[d] = gpu.device_alloc(grids[0].nbytes)
    for g, (x, y) in zip(grids, poses):
        if devices is None:
            gpu.memcpy_h2d(dev, g)
            ptrs = [dev + bd.row0[r] * bd.G * 16 for r in range(bands)]
        else:
            for d, p in per_dev.items():
                gpu.set_device(d)
                gpu.memcpy_h2d(p, g)
            ptrs = [per_dev[devices[r]] + bd.row0[r] * bd.G * 16 for r in range(bands)]
        history.append(bd.update_grid(ptrs, x, y, 0.0, 0.1))
    gpu.set_device(0)
    gpu.device_free(dev)
    bd._extra_buffers = per_dev  # (freed with the process; kept alive while the bands exist)
    return bd, history


def poses(steps, vy=4.0, vx=0.0):
    return [(np.float32(vx * 0.1 * (s + 1)), np.float32(vy * 0.1 * (s + 1))) for s in range(steps)]


def test_one_band_reproduces_the_single_grid_cycle(gpu):
    grids, ps = scans(gpu, 6), poses(6)
    plain = run_plain(gpu, grids, ps)
    bd, hist = run_banded(gpu, grids, ps, 1)
    assert all(h == [N] for h in hist)
    state, idx, w, a = bd.get_particles(0)
    ref = plain.get_particles()
    assert np.array_equal(state.view(np.uint32), ref.state.view(np.uint32))
    assert np.array_equal(idx, ref.grid_cell_idx) and np.array_equal(w.view(np.uint32), ref.weight.view(np.uint32))
    assert np.array_equal(bd.get_grid_cells().view(np.uint8), plain.get_grid_cells().view(np.uint8))
    bd.close()
    plain.close()


@pytest.mark.parametrize("bands,vx,vy", [(2, 0.0, 4.0), (3, 3.0, 0.0), (4, -2.0, 6.0)])
def test_bands_conserve_and_agree_with_the_single_grid(gpu, bands, vx, vy):
    steps = 8
    grids, ps = scans(gpu, steps), poses(steps, vy, vx)
    plain = run_plain(gpu, grids, ps)
    bd, hist = run_banded(gpu, grids, ps, bands)
    G = bd.G
    # every cycle hands out exactly the global number of particles, or one fewer per band edge through float rounding
    for counts in hist:
        assert N - bands <= sum(counts) <= N
    total_w, n_seen = 0.0, 0
    for r in range(bands):
        state, idx, w, _ = bd.get_particles(r)
        n_seen += len(w)
        total_w += float(w.astype(np.float64).sum())
        # the resampled particles are copies of particles (or birth particles) that were in the band
        assert np.all(idx >= 0) and np.all(idx < bd.rows[r] * G)
        rows = state[:, 1].astype(np.int32)
        inside = (state[:, 1] >= 0) & (state[:, 1] <= G - 1)
        # (birth particles sit at y = cell / float(G) + 0.5, init_new_particles.cu:173, i.e. up to half a cell above
        #  their cell's row: those of a band's last row may already lie in the next band and move there next cycle)
        assert np.all((rows[inside] >= bd.row0[r]) & (rows[inside] <= bd.row0[r] + bd.rows[r]))
        assert np.unique(w).size == 1  # weight = joint total / N of the whole grid
    assert n_seen == sum(hist[-1])
    pw = plain.get_particles().weight.astype(np.float64).sum()
    assert abs(total_w - pw) <= 0.02 * pw
    # particles did cross band edges (the test would be vacuous otherwise)
    lo, hi = bd.last_migration
    assert sum(lo) + sum(hi) > 0
    # the maps agree statistically: occupied mass in 10 x 10 blocks
    cb, cp = bd.get_grid_cells(), plain.get_grid_cells()
    assert cb.size == cp.size == G * G

    def blocks(cells, field):
        v = cells[field].astype(np.float64).reshape(G, G)
        return v.reshape(G // 10, 10, G // 10, 10).sum(axis=(1, 3))

    # yardstick: the same single-grid run with another seed (two independent realisations of the same filter)
    other = run_plain(gpu, grids, ps, seed=78)
    co = other.get_grid_cells()
    other.close()
    ob, op, oo = blocks(cb, "occ_mass"), blocks(cp, "occ_mass"), blocks(co, "occ_mass")
    big = op > 2.0
    rel = np.abs(ob[big] - op[big]) / op[big]
    rel_seed = np.abs(oo[big] - op[big]) / op[big]
    print("blocks", int(big.sum()), "banded vs single: median", np.median(rel), "p90", np.percentile(rel, 90),
          "| seed vs seed: median", np.median(rel_seed), "p90", np.percentile(rel_seed, 90))
    assert big.sum() > 20
    assert np.median(rel) < 1.5 * np.median(rel_seed) + 0.005 and np.percentile(rel, 90) < 1.5 * np.percentile(rel_seed, 90) + 0.01
    fb, fp = blocks(cb, "free_mass"), blocks(cp, "free_mass")
    assert np.allclose(fb, fp, rtol=0.02, atol=0.3)
    # the velocity estimate of well-populated cells agrees too
    sel = (cp["occ_mass"] > 0.6) & (cb["occ_mass"] > 0.6)
    assert sel.sum() > 50
    dv = np.abs(cb["mean_y_vel"][sel] - cp["mean_y_vel"][sel])
    assert np.median(dv) < 6.0  # cells / s, i.e. 1.2 m/s
    bd.close()
    plain.close()


@pytest.mark.parametrize("bands", [2, 3])
def test_bands_are_deterministic(gpu, bands):
    """Two runs with the same seed end in bit-identical particles and maps: the particles that cross a band edge reach the
    neighbour in slot order (k_outbox_*), not in the order of an atomic append."""
    grids, ps = scans(gpu, 7), poses(7, vy=5.0, vx=1.0)
    results = []
    for _ in range(2):
        bd, hist = run_banded(gpu, grids, ps, bands)
        parts = [tuple(np.ascontiguousarray(a).copy() for a in bd.get_particles(r)) for r in range(bands)]
        results.append((hist, parts, bd.get_grid_cells().copy()))
        bd.close()
    (h0, p0, g0), (h1, p1, g1) = results
    assert h0 == h1
    assert sum(h0[-1]) == N and min(h0[-1]) > 0
    for r in range(bands):
        for a, b in zip(p0[r], p1[r]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"band {r}"
    assert np.array_equal(g0.view(np.uint8), g1.view(np.uint8))


def test_native_band_group_equals_the_phase_calls(gpu):
    """dogm_band_group_* (host threads inside the library) runs the same phases as the dogm_band_* calls driven from Python:
    bit-identical particles and maps."""
    grids, ps = scans(gpu, 6), poses(6, vy=5.0, vx=-1.0)
    out = []
    for native in (True, False):
        bd, hist = run_banded(gpu, grids, ps, 3, native=native)
        assert (bd.group is not None) == native
        parts = [tuple(np.ascontiguousarray(a).copy() for a in bd.get_particles(r)) for r in range(3)]
        out.append((hist, parts, bd.get_grid_cells().copy(), dict(bd.last_totals)))
        bd.close()
    assert out[0][0] == out[1][0] and out[0][3] == out[1][3]
    for r in range(3):
        for a, b in zip(out[0][1][r], out[1][1][r]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"band {r}"
    assert np.array_equal(out[0][2].view(np.uint8), out[1][2].view(np.uint8))


def _snapshot(bd, bands):
    parts = [tuple(np.ascontiguousarray(a).copy() for a in bd.get_particles(r)) for r in range(bands)]
    return parts, bd.get_grid_cells().copy(), dict(bd.last_totals)


@pytest.mark.parametrize("bands,vx,vy", [(1, 0.0, 4.0), (2, 0.0, 4.0), (3, 3.0, 0.0), (4, -2.0, 6.0)])
def test_device_paced_cycle_equals_the_host_paced_phases(gpu, bands, vx, vy):
    """The device-paced group (whole cycle enqueued at once; counts, records, halo rows and normalisers exchanged GPU to GPU by
    the kernels themselves) and the host-paced phases end in bit-identical particles, maps, counts and normalisers."""
    grids, ps = scans(gpu, 7), poses(7, vy, vx)
    out = []
    for paced in (True, False):
        bd, hist = run_banded(gpu, grids, ps, bands, device_paced=paced)
        assert bd.device_paced == paced
        out.append((hist,) + _snapshot(bd, bands))
        if bands > 1:
            lo, hi = bd.last_migration
            assert sum(lo) + sum(hi) > 0
        bd.close()
    assert out[0][0] == out[1][0], (out[0][0], out[1][0])
    assert out[0][3] == out[1][3]
    for r in range(bands):
        for a, b in zip(out[0][1][r], out[1][1][r]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"band {r}"
    assert np.array_equal(out[0][2].view(np.uint8), out[1][2].view(np.uint8))


def test_device_paced_launch_estimates_do_not_matter(gpu, monkeypatch):
    """The device-paced cycle sizes its launches from estimates (the counts live on the device) and the kernels loop: with
    absurdly small estimates the result is the same bit for bit."""
    grids, ps = scans(gpu, 6), poses(6, vy=5.0, vx=1.0)
    out = []
    for est in (None, "3000"):
        if est:
            monkeypatch.setenv("DOGM_B200_BAND_EST", est)
        bd, hist = run_banded(gpu, grids, ps, 3, device_paced=True)
        out.append((hist,) + _snapshot(bd, 3))
        bd.close()
    monkeypatch.delenv("DOGM_B200_BAND_EST", raising=False)
    assert out[0][0] == out[1][0] and out[0][3] == out[1][3]
    for r in range(3):
        for a, b in zip(out[0][1][r], out[1][1][r]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"band {r}"
    assert np.array_equal(out[0][2].view(np.uint8), out[1][2].view(np.uint8))


def test_bands_on_two_gpus_equal_bands_on_one(gpu):
    """Two bands on two GPUs (records, halo rows and normalisers cross NVLink: peer loads and stores from inside the kernels)
    end in the same bits as the same two bands on one GPU."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    grids, ps = scans(gpu, 7), poses(7, vy=5.0, vx=1.0)
    out = []
    for devices in ([0, 1], [0, 0]):
        bd, hist = run_banded(gpu, grids, ps, 2, device_paced=True, devices=devices)
        assert bd.device_paced
        out.append((hist,) + _snapshot(bd, 2))
        bd.close()
    assert out[0][0] == out[1][0]
    assert out[0][3] == out[1][3]
    for r in range(2):
        for a, b in zip(out[0][1][r], out[1][1][r]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"band {r}"
    assert np.array_equal(out[0][2].view(np.uint8), out[1][2].view(np.uint8))


@pytest.mark.parametrize("bands,paced", [(2, True), (4, True), (3, False)])
def test_bands_with_injected_state_reproduce_the_single_grid_masses(gpu, bands, paced):
    """An oracle for band mode: the single grid.  Both start one cycle from the SAME state (a realistic population and map
    taken from a running filter, loaded through dogm_band_set_state / dogm_set_*), with the process noise switched off so
    that the prediction is the same function on both sides.  The bands then hold the same particles per cell as the single
    grid (some arrived over a band edge, in another order): every per-cell quantity that is an order-independent sum - the
    predicted, updated, born and persistent masses, the born-mass array, the cell ranges' lengths - must be the same bits,
    the joint weight total the same to double rounding, and the global particle count must be conserved."""
    grids, ps = scans(gpu, 5), poses(5, vy=5.0, vx=2.0)
    warm = run_plain(gpu, grids[:4], ps[:4])
    P, G = warm.get_particles(), warm.get_grid_cells()
    pose = (warm.get_position_x(), warm.get_position_y(), warm.get_yaw())
    warm.close()
    quiet = make_params(gpu, SIZE, RES, N, B, stddev_process_noise_position=0.0, stddev_process_noise_velocity=0.0)
    (x, y) = ps[4]
    # single grid
    d = gpu.DOGM(quiet)
    d.set_options(seed=5, resample_mode=gpu.RESAMPLE_SYSTEMATIC, noise_mode=gpu.NOISE_PHILOX)
    d.update_measurement_grid(grids[3])  # (the first measurement initialises a population; it is replaced right away)
    d.set_particles(P)
    d.set_grid_cells(G)
    d.set_pose(*pose)
    d.update_grid(grids[4], x, y, 0.0, 0.1, device=False)
    cells_single, born_single = d.get_grid_cells(), d.get_born_masses()
    total_single = d.get_joint_weight_accum()[-1]
    d.close()
    # the bands
    bd = gpu.BandedDOGM(quiet, bands, seed=5, device_paced=paced)
    bd.set_state(P.state, P.weight, P.associated, G, *pose)
    dev = gpu.device_alloc(grids[4].nbytes)
    gpu.memcpy_h2d(dev, grids[4])
    counts = bd.update_grid([dev + bd.row0[r] * bd.G * 16 for r in range(bands)], x, y, 0.0, 0.1)
    cells_band = bd.get_grid_cells()
    lo, hi = bd.last_migration
    assert sum(lo) + sum(hi) > 0, "no particle crossed a band edge: the test would be vacuous"
    assert N - bands <= sum(counts) <= N
    for f in ("pred_occ_mass", "occ_mass", "free_mass", "new_born_occ_mass", "pers_occ_mass"):
        assert np.array_equal(cells_band[f].view(np.uint32), cells_single[f].view(np.uint32)), f
    n_single = np.where(cells_single["start_idx"] >= 0, cells_single["end_idx"] - cells_single["start_idx"] + 1, 0)
    n_band = np.where(cells_band["start_idx"] >= 0, cells_band["end_idx"] - cells_band["start_idx"] + 1, 0)
    # (a particle that left its band stays behind as a weightless ghost in the band's edge row until resampling drops it, so the
    #  edge rows may count more particles than the single grid; a particle that left the GRID sideways has weight 0, sits in
    #  the first or last column and is not sent on when it also crosses a band edge; everywhere else the counts are the same)
    inner = np.ones((bd.G, bd.G), bool)
    inner[:, 0] = inner[:, -1] = False
    for r in range(bands):
        inner[bd.row0[r]] = False
        inner[bd.row0[r] + bd.rows[r] - 1] = False
    inner = inner.ravel()
    assert np.array_equal(n_single[inner], n_band[inner]), "the bands hold other particles per cell than the single grid"
    # every migrated particle is counted twice in the bands' cell ranges: as the ghost it left behind and where it arrived
    assert int(n_single.sum()) == N and int(n_band.sum()) == N + sum(lo) + sum(hi)
    assert abs(bd.last_totals["weight"] - total_single) <= 1e-11 * total_single
    born_total_single = float(born_single.astype(np.float64).sum())
    assert abs(bd.last_totals["born"] - born_total_single) <= 1e-9 * max(born_total_single, 1.0)
    # mean velocities: float sums in another order inside a cell
    sel = cells_single["pers_occ_mass"] > 1e-3
    assert np.allclose(cells_band["mean_x_vel"][sel], cells_single["mean_x_vel"][sel], rtol=1e-4, atol=1e-3)
    assert np.allclose(cells_band["mean_y_vel"][sel], cells_single["mean_y_vel"][sel], rtol=1e-4, atol=1e-3)
    gpu.device_free(dev)
    bd.close()
