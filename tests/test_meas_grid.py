"""Measurement grid from a lidar scan: CPU checks of the oracle's restatement, GPU checks of the CUDA kernel against
the oracle and (stage M1, the polar inverse sensor model) against the reference's own kernel."""
import numpy as np
import pytest

from _loader import load_ref

DEMO_LASER = dict(max_range=50.0, resolution=0.2, fov=120.0, stddev_range=0.5)  # demo/main.cpp:36-41


def demo_beams(rng, k=100, max_range=50.0):
    beams = np.full(k, np.inf, np.float32)
    hit = rng.uniform(size=k) < 0.6
    beams[hit] = rng.uniform(2.0, max_range * 0.95, size=int(hit.sum())).astype(np.float32)
    return beams


def polar_numpy(beams, H, res, sigma):
    """Independent float64 statement of measurement_grid.cu:33-89."""
    K = beams.size
    out = np.zeros((H, K, 2))
    for b in range(K):
        z = float(beams[b])
        for i in range(H):
            free_p = 0.15 + i * 0.85 / H
            if np.isfinite(z):
                r = int(np.float32(z) / np.float32(res))
                occ = 0.95 * np.exp(-0.5 * ((i - r) * res) ** 2 / sigma**2)
                if i <= r:
                    m = (occ, 0.0) if occ > free_p else (0.0, 1.0 - free_p)
                else:
                    m = (occ, 0.0) if occ > 0.5 else (0.0, 0.0)
            else:
                m = (0.0, 1.0 - free_p)
            out[i, b] = np.clip(m, 1e-5, 1 - 1e-5)
    return out


def test_oracle_polar_model_matches_numpy(orc):
    rng = np.random.default_rng(41)
    laser = orc.LaserParams(10.0, 0.2, 120.0, 0.5)
    beams = demo_beams(rng, 37, 10.0)
    got = orc.meas_polar_grid(laser, beams)
    exp = polar_numpy(beams, 50, 0.2, 0.5)
    # thresholds (occ > free_p, occ > 0.5) may flip within float rounding of expf: allow isolated texels
    bad = np.abs(got - exp) > 1e-5
    assert bad.sum() <= 2, bad.sum()


def test_oracle_warp_geometry(orc):
    """Sensor at the bottom centre of the rendered image = last row of the measurement grid (row flip of
    measurement_grid.cu:120); nothing outside the field of view; free space along beams without return."""
    laser = orc.LaserParams(50.0, 0.2, 120.0, 0.5)
    beams = np.full(100, np.inf, np.float32)
    m = orc.meas_generate(laser, 50.0, 0.2, beams).reshape(250, 250)
    assert np.all(m["likelihood"] == 1.0) and np.all(m["p_A"] == 1.0)
    # the fan opens upwards in the image = towards smaller grid rows; straight ahead is free, the bottom corners are not seen
    assert m["free_mass"][125, 125] > 0.2 and m["occ_mass"][125, 125] <= 1.01e-5
    assert m["free_mass"][249, 5] == 0.0 and m["occ_mass"][249, 5] == 0.0
    assert m["free_mass"][245, 245] == 0.0
    # free mass decays with range (pFree rises from 0.15 to 1)
    col = m["free_mass"][:, 125]
    assert col[240] > col[120] > col[10] > 0
    # returns straight ahead at 20 m (two neighbouring beams, so that the bilinear fetch in between sees only them):
    # an occupied arc about 100 cells above the sensor row
    beams[49] = beams[50] = 20.0
    m2 = orc.meas_generate(laser, 50.0, 0.2, beams).reshape(250, 250)
    occ_rows = np.nonzero(m2["occ_mass"][:, 125:126].max(axis=1) > 0.5)[0]
    assert occ_rows.size > 0 and abs(float(occ_rows.mean()) - (249 - 100)) < 4
    # behind the return nothing is known
    assert m2["free_mass"][100, 125] <= 1.01e-5 and m2["occ_mass"][100, 125] <= 1.01e-5


@pytest.mark.gpu
@pytest.mark.parametrize("grid_length,res,k", [(50.0, 0.2, 100), (12.0, 0.1, 48), (30.0, 0.25, 77)])
def test_cuda_measurement_grid_matches_oracle(gpu, orc, grid_length, res, k):
    rng = np.random.default_rng(42)
    lp = dict(DEMO_LASER, max_range=grid_length, resolution=res)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(lp["max_range"], lp["resolution"], lp["fov"], lp["stddev_range"]), grid_length, res)
    laser = orc.LaserParams(lp["max_range"], lp["resolution"], lp["fov"], lp["stddev_range"])
    for _ in range(2):
        beams = demo_beams(rng, k, grid_length)
        got = gen.generate_grid_host(beams)
        exp = orc.meas_generate(laser, grid_length, res, beams)
        assert got.size == exp.size
        for f in ("occ_mass", "free_mass"):
            bad = np.abs(got[f] - exp[f]) > 2e-5
            # expf / atan2f differ by an ulp between libm and CUDA: a threshold of the sensor model can flip in isolated texels
            assert bad.mean() < 2e-4, (f, int(bad.sum()))
        assert np.all(got["likelihood"] == 1.0) and np.all(got["p_A"] == 1.0)
        pg = gen.polar_grid(beams)
        pe = orc.meas_polar_grid(laser, beams)
        assert (np.abs(pg - pe) > 1e-5).sum() <= 4
    gen.close()


@pytest.mark.gpu
def test_cuda_polar_model_matches_reference_kernel(gpu):
    ref = load_ref()
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(43)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(50.0, 0.2, 120.0, 0.5), 50.0, 0.2)
    beams = demo_beams(rng, 100, 50.0)
    mine = gen.polar_grid(beams)
    theirs = ref.polar_grid(beams, 250, 0.2, 0.5)  # createPolarGridTextureKernel on a CUDA surface
    assert mine.shape == theirs.shape
    assert (np.abs(mine - theirs) > 1e-6).sum() <= 2  # same expf on the same device; FMA contraction may flip a threshold


@pytest.mark.gpu
def test_generate_into_feeds_update_grid(gpu):
    from conftest import make_params

    rng = np.random.default_rng(44)
    p = make_params(gpu, 50.0, 0.2, 50000, 5000)
    d = gpu.DOGM(p)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(50.0, 0.2, 120.0, 0.5), 50.0, 0.2)
    beams = demo_beams(rng)
    host = gen.generate_grid_host(beams)
    gen.generate_grid_into(d, beams)
    d.update_grid(None, 0.0, 0.0, 0.0, 0.1)
    assert np.array_equal(d.get_measurement_cells().view(np.uint8), host.view(np.uint8))
    e = gpu.DOGM(p)
    ptr = gen.generate_grid(beams)
    e.update_grid(ptr, 0.0, 0.0, 0.0, 0.1, device=True)
    assert np.array_equal(d.get_grid_cells().view(np.uint8), e.get_grid_cells().view(np.uint8))
    assert np.array_equal(d.get_particles().block, e.get_particles().block)
    # later cycles: the cartesian resampling runs inside the cycle's cell kernel (no measurement-grid pass of its own);
    # the handle's measurement grid, the map and the particles equal those of the two-step path, with grid shifts
    for c in range(1, 5):
        beams = demo_beams(rng)
        host = gen.generate_grid_host(beams)
        ptr = gen.generate_grid(beams)
        e.update_grid(ptr, 0.3 * c, 0.5 * c, 0.0, 0.1, device=True)
        gen.generate_grid_into(d, beams)
        if c == 2:  # looking at the grid before the cycle forces it into existence; the cycle then takes the plain path
            assert np.array_equal(d.get_measurement_cells().view(np.uint8), host.view(np.uint8))
        d.update_grid(None, 0.3 * c, 0.5 * c, 0.0, 0.1, sync=(c != 3))
        assert np.array_equal(d.get_measurement_cells().view(np.uint8), host.view(np.uint8)), c
        assert np.array_equal(d.get_grid_cells().view(np.uint8), e.get_grid_cells().view(np.uint8)), c
        assert np.array_equal(d.get_particles().block, e.get_particles().block), c
    # a scan that is generated but superseded by a caller-supplied grid is dropped
    other = demo_beams(rng)
    gen.generate_grid_into(d, other)
    beams = demo_beams(rng)
    host = gen.generate_grid_host(beams)
    ptr = gen.generate_grid(beams)
    d.update_grid(ptr, 1.5, 2.5, 0.0, 0.1, device=True)
    e.update_grid(ptr, 1.5, 2.5, 0.0, 0.1, device=True)
    assert np.array_equal(d.get_measurement_cells().view(np.uint8), host.view(np.uint8))
    assert np.array_equal(d.get_grid_cells().view(np.uint8), e.get_grid_cells().view(np.uint8))
    d.update_grid(None, 1.5, 2.5, 0.0, 0.1)  # and does not come back
    assert np.array_equal(d.get_measurement_cells().view(np.uint8), host.view(np.uint8))


@pytest.mark.gpu
def test_generate_into_many_beams(gpu):
    """More beams than travel in the kernel parameters (960): the scan goes through the device copy; same grid either way."""
    from conftest import make_params

    rng = np.random.default_rng(45)
    p = make_params(gpu, 30.0, 0.1, 20000, 2000)
    d = gpu.DOGM(p)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(30.0, 0.1, 120.0, 0.5), 30.0, 0.1)
    for k in (1200, 960, 961, 7):
        for c in range(2):
            beams = demo_beams(rng, k, 30.0)
            host = gen.generate_grid_host(beams)
            gen.generate_grid_into(d, beams)
            d.update_grid(None, 0.0, 0.2 * c, 0.0, 0.1)
            assert np.array_equal(d.get_measurement_cells().view(np.uint8), host.view(np.uint8)), (k, c)


def test_oracle_fusion_properties(orc):
    """combine_masses (measurement_grid.cu:13-31): fusing a scan without information changes nothing much, fusing the same
    evidence twice sharpens it, masses stay in [0, 1] and their sum stays <= 1."""
    laser = orc.LaserParams(50.0, 0.2, 120.0, 0.5)
    rng = np.random.default_rng(45)
    a = demo_beams(rng, 100, 50.0)
    once = orc.meas_polar_fused(laser, a[None, :])
    assert np.array_equal(once, orc.meas_polar_grid(laser, a))
    twice = orc.meas_polar_fused(laser, np.stack([a, a]))
    assert np.all(twice >= -1e-6) and np.all(twice.sum(axis=2) <= 1.0 + 1e-5)
    occ_cells = once[..., 0] > 0.5
    assert occ_cells.any() and np.all(twice[..., 0][occ_cells] >= once[..., 0][occ_cells] - 1e-6)
    free_cells = once[..., 1] > 0.2
    assert np.all(twice[..., 1][free_cells] >= once[..., 1][free_cells] - 1e-6)
    grid = orc.meas_generate_fused(laser, 50.0, 0.2, np.stack([a, a]))
    assert grid.size == 250 * 250 and np.all(grid["likelihood"] == 1.0)


@pytest.mark.gpu
def test_cuda_fused_scans_match_oracle_and_reference_kernel(gpu, orc):
    """dogm_meas_generate_fused: the fused polar grid against the C restatement and against the reference's own
    createPolarGridTextureKernel + fusePolarGridTextureKernel run on a CUDA surface; the cartesian grid against the oracle."""
    rng = np.random.default_rng(46)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(50.0, 0.2, 120.0, 0.5), 50.0, 0.2)
    laser = orc.LaserParams(50.0, 0.2, 120.0, 0.5)
    scans = np.stack([demo_beams(rng, 100, 50.0) for _ in range(3)])
    ptr, polar = gen.generate_grid_fused(scans, want_polar=True)
    pe = orc.meas_polar_fused(laser, scans)
    assert polar.shape == pe.shape and (np.abs(polar - pe) > 2e-5).sum() <= 6
    got = np.empty(250 * 250, gpu.MEAS_CELL_DTYPE)
    gpu.memcpy_d2h(got, ptr)
    exp = orc.meas_generate_fused(laser, 50.0, 0.2, scans)
    for f in ("occ_mass", "free_mass"):
        assert (np.abs(got[f] - exp[f]) > 5e-5).mean() < 3e-4, f
    # one scan through the fused entry equals the plain entry
    one = np.empty(250 * 250, gpu.MEAS_CELL_DTYPE)
    gpu.memcpy_d2h(one, gen.generate_grid_fused(scans[:1]))
    assert np.array_equal(one.view(np.uint8), gen.generate_grid_host(scans[0]).view(np.uint8))
    ref = load_ref()
    if ref.available() and hasattr(ref.load_library(), "ref_polar_fused"):
        theirs = ref.polar_fused(scans, 250, 0.2, 0.5)
        assert (np.abs(polar - theirs) > 2e-6).sum() <= 6  # same device, same expf; FMA contraction on their side only
    gen.close()


@pytest.mark.gpu
@pytest.mark.parametrize("readout", [False, True])
def test_generate_into_back_to_back_on_a_small_grid(gpu, readout):
    """Scans and cycles enqueued back to back without waiting (a small grid: the CTAs of a whole cycle fit on the device at
    once, each kernel waiting for the one in front): the polar table of scan c+1 must not overtake the cell kernel of cycle c
    that still reads the table of scan c.  Same final state as the blocking two-step path."""
    from conftest import make_params

    rng = np.random.default_rng(46)
    p = make_params(gpu, 20.0, 0.2, 30000, 3000)
    d, e = gpu.DOGM(p), gpu.DOGM(p)
    gen = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(20.0, 0.2, 120.0, 0.5), 20.0, 0.2)
    gen2 = gpu.LaserMeasurementGrid(gpu.LaserSensorParams(20.0, 0.2, 120.0, 0.5), 20.0, 0.2)
    scans = [demo_beams(rng, 60, 20.0) for _ in range(25)]
    if readout:
        d.set_dynamic_cell_filter(0.6, 0.5, 4096)
    for c, z in enumerate(scans):
        gen.generate_grid_into(d, z)
        d.update_grid(None, 0.1 * c, 0.3 * c, 0.0, 0.1, sync=False)
        if readout:
            d.extract_dynamic_cells(0.6, 0.5, capacity=4096)
    for c, z in enumerate(scans):
        e.update_grid(gen2.generate_grid(z), 0.1 * c, 0.3 * c, 0.0, 0.1, device=True)
    assert np.array_equal(d.get_grid_cells().view(np.uint8), e.get_grid_cells().view(np.uint8))
    assert np.array_equal(d.get_particles().block, e.get_particles().block)
    assert np.array_equal(d.get_measurement_cells().view(np.uint8), e.get_measurement_cells().view(np.uint8))
