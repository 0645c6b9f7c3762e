"""Host-side companions: the lidar scene simulator, DBSCAN and the MAE / RMSE evaluator in
include/*.h, driven through libdogm_b200_tools.so.
  * the reference's simulator_spec.cpp (dogm/demo/simulator/test/simulator_spec.cpp:24-52) restated;
  * bit-exact comparison with the reference's own classes compiled from /root/reference (oracle/_ref/libdemo_ref.so, built by
    oracle/build_demo_ref.sh, CPU only) on the demo scenes and on random inputs, when that library is present;
  * the committed golden fixture tests/golden/demo_tools.npz (recorded from the reference's classes) otherwise and always."""
import importlib.util
import os

import numpy as np
import pytest

from _loader import PKG_DIR, ROOT

spec = importlib.util.spec_from_file_location("dogm_b200_tools", os.path.join(PKG_DIR, "tools.py"))
tools_mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tools_mod)

REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libdemo_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "demo_tools.npz")
DEMO = dict(num_points=100, fov=120.0, grid_size=50.0, ego_velocity=(0.0, 4.0), steps=14, dt=0.1)  # demo/main.cpp:36-48


@pytest.fixture(scope="module")
def mine():
    return tools_mod.Tools()


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libdemo_ref.so not built (needs /root/reference)")
    return tools_mod.Tools(REF_LIB, "ref_tools_")


def random_cells(rng, n, gs):
    idx = rng.choice(gs * gs, size=n, replace=False).astype(np.int32)
    cells = np.zeros(n, tools_mod.DYNAMIC_CELL_DTYPE)
    cells["cell_idx"] = idx
    cells["mean_x_vel"] = rng.normal(0, 40, n)
    cells["mean_y_vel"] = rng.normal(0, 40, n)
    return cells


def clustered_cells(rng, gs, centres, per=40, spread=2.0):
    pts = []
    for cx, cy, vx, vy in centres:
        x = np.clip(np.round(rng.normal(cx, spread, per)), 0, gs - 1).astype(np.int64)
        y = np.clip(np.round(rng.normal(cy, spread, per)), 0, gs - 1).astype(np.int64)
        for a, b in zip(x, y):
            pts.append((int(b * gs + a), vx + rng.normal(0, 2), vy + rng.normal(0, 2)))
    uniq = {}
    for i, vx, vy in pts:
        uniq.setdefault(i, (vx, vy))
    cells = np.zeros(len(uniq), tools_mod.DYNAMIC_CELL_DTYPE)
    cells["cell_idx"] = list(uniq.keys())
    cells["mean_x_vel"] = [v[0] for v in uniq.values()]
    cells["mean_y_vel"] = [v[1] for v in uniq.values()]
    rng.shuffle(cells)
    return cells


def test_all_entry_points_are_exported():
    import re
    import subprocess

    text = open(os.path.join(ROOT, "include", "dogm_b200_tools.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(dogm_tools_[a-z0-9_]+)\s*\(", text)))
    assert declared == sorted("dogm_tools_" + n for n in tools_mod.ENTRY_POINTS)
    out = subprocess.run(["nm", "-D", "--defined-only", tools_mod.LIB_PATH], capture_output=True, text=True).stdout
    assert set(declared) <= set(re.findall(r"\bT (dogm_tools_[a-z0-9_]+)", out))


def test_vehicle_spec(mine):
    """simulator_spec.cpp: Constructor / GetFacingSide / Move"""
    width, pos, vel = 3.5, (np.float32(-12.3), np.float32(20.2)), (5.0, -10.0)
    pts, count = mine.facing_side((width, pos[0], pos[1], vel[0], vel[1]), 1.5)
    left = np.float32(pos[0]) + np.float32(-width * 0.5)
    expected = np.array([[left, pos[1]], [left + np.float32(1.5), pos[1]], [left + np.float32(2.0) * np.float32(1.5), pos[1]]], np.float32)
    assert count == 3 and np.array_equal(pts, expected)
    for dt in (-3.2, 0.0, 3.4, 10.5):  # one simulator step with a resting ego vehicle = Vehicle::move(dt)
        _, states, ego = mine.simulate(8, 120.0, 50.0, (0.0, 0.0), [(width, pos[0], pos[1], vel[0], vel[1])], 1, dt)
        exp = np.array([pos[0] + np.float32(dt) * np.float32(vel[0]), pos[1] + np.float32(dt) * np.float32(vel[1])], np.float32)
        assert np.array_equal(states[0, 0, :2], exp) and np.array_equal(states[0, 0, 2:], np.float32(vel))
        assert np.array_equal(ego[0], [0.0, 0.0])


def test_demo_scene_properties(mine):
    meas, states, ego = mine.simulate(vehicles=tools_mod.DEMO_VEHICLES, **DEMO)
    assert meas.shape == (14, 100) and states.shape == (14, 4, 4)
    assert np.allclose(ego[:, 1], 0.4 * np.arange(1, 15), rtol=1e-6) and np.all(ego[:, 0] == 0)
    hit = np.isfinite(meas)
    assert hit.any(axis=1).all() and (~hit).any(axis=1).all()
    assert np.all(meas[hit] > 0) and np.all(meas[hit] < 50.0 * 1.5)
    # the resting vehicle (45, 15) drifts towards the sensor with the ego motion only
    assert np.allclose(states[:, 3, 1], 15.0 - 0.4 * np.arange(1, 15), atol=1e-4)


def test_simulator_matches_reference(mine, ref):
    rng = np.random.default_rng(3)
    scenes = [(DEMO, tools_mod.DEMO_VEHICLES), (DEMO, tools_mod.DEMO_VEHICLES_ALT)]
    for _ in range(6):
        gs = float(rng.choice([50.0, 120.0, 409.6]))
        cfg = dict(num_points=int(rng.integers(20, 600)), fov=float(rng.choice([60.0, 90.0, 120.0, 150.0, 180.0])), grid_size=gs,
                   ego_velocity=(float(rng.uniform(-3, 3)), float(rng.uniform(0, 10))), steps=int(rng.integers(1, 20)), dt=0.1)
        veh = [(rng.uniform(1.5, 6), rng.uniform(0, gs), rng.uniform(1, gs), rng.uniform(-15, 15), rng.uniform(-15, 15)) for _ in range(int(rng.integers(0, 9)))]
        scenes.append((cfg, veh))
    for cfg, veh in scenes:
        a, b = mine.simulate(vehicles=veh, **cfg), ref.simulate(vehicles=veh, **cfg)
        for x, y, what in zip(a, b, ("measurements", "vehicle states", "ego pose")):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), what
    for v in [(3.5, -12.3, 20.2, 5, -10), (0.01, 3.0, 4.0, 0, 0), (6.0, 100.0, 1.0, 0, 0)]:
        for res in (1.5, 0.025, 0.3):
            pa, ca = mine.facing_side(v, res)
            pb, cb = ref.facing_side(v, res)
            assert ca == cb and np.array_equal(pa.view(np.uint32), pb.view(np.uint32))


def test_dbscan_matches_reference(mine, ref):
    rng = np.random.default_rng(4)
    for case in range(12):
        gs = 250
        if case % 2:
            cells = clustered_cells(rng, gs, [(rng.uniform(20, 230), rng.uniform(20, 230), 0, 0) for _ in range(int(rng.integers(1, 6)))],
                                    per=int(rng.integers(3, 60)), spread=float(rng.uniform(0.5, 4)))
        else:
            cells = random_cells(rng, int(rng.integers(0, 300)), int(rng.integers(12, 60)))
            gs = 60
        cells = np.sort(cells, order="cell_idx")
        xy = np.stack([cells["cell_idx"] % gs, cells["cell_idx"] // gs], axis=1).astype(np.float32)
        for eps, mc in ((3.0, 5), (1.5, 3), (6.0, 12)):
            la, ca = mine.dbscan(xy, eps, mc)
            lb, cb = ref.dbscan(xy, eps, mc)
            assert ca == cb and np.array_equal(la, lb), (case, eps, mc)


def evaluate(t, cells_per_step, gs, vehicles, resolution=0.2):
    ev = t.evaluator(vehicles=vehicles, resolution=resolution, **DEMO)
    for s, cells in enumerate(cells_per_step):
        ev.step(s, cells, gs)
    out = ev.summary()
    ev.close()
    return out


def demo_like_cells(rng, t, gs=250, res=0.2):
    """dynamic cells around the simulated vehicles (grid y axis points down: row = gs - y / res), plus clutter"""
    _, states, _ = t.simulate(vehicles=tools_mod.DEMO_VEHICLES, **DEMO)
    steps = []
    for s in range(states.shape[0]):
        centres = [(v[0] / res, gs - v[1] / res, v[2] / res + rng.normal(0, 3), -v[3] / res + rng.normal(0, 3)) for v in states[s]]
        cells = clustered_cells(rng, gs, centres, per=30, spread=1.5)
        clutter = random_cells(rng, 25, gs)
        clutter = clutter[~np.isin(clutter["cell_idx"], cells["cell_idx"])]
        steps.append(np.concatenate([cells, clutter]))
    return steps


def test_evaluator_matches_reference(mine, ref):
    rng = np.random.default_rng(5)
    steps = demo_like_cells(rng, mine)
    a, b = evaluate(mine, steps, 250, tools_mod.DEMO_VEHICLES), evaluate(ref, steps, 250, tools_mod.DEMO_VEHICLES)
    assert a["detections"] == b["detections"] > 20 and a["unassigned"] == b["unassigned"]
    assert np.array_equal(a["mae"].view(np.uint32), b["mae"].view(np.uint32))
    assert np.array_equal(a["rmse"].view(np.uint32), b["rmse"].view(np.uint32))
    assert np.all(a["mae"][:2] < 1.0) and np.all(a["mae"][2:] < 2.0)  # the synthetic cells do sit on the vehicles
    # empty steps and steps without any cluster are fine
    empty = [np.zeros(0, tools_mod.DYNAMIC_CELL_DTYPE)] * 14
    ea, eb = evaluate(mine, empty, 250, tools_mod.DEMO_VEHICLES), evaluate(ref, empty, 250, tools_mod.DEMO_VEHICLES)
    assert ea["detections"] == eb["detections"] == 0 and ea["unassigned"] == eb["unassigned"] == 0


def test_against_golden_fixture(mine):
    """tests/golden/demo_tools.npz was recorded from the reference's classes by tests/golden/make_demo_tools_golden.py"""
    g = np.load(GOLDEN)
    meas, states, ego = mine.simulate(vehicles=tools_mod.DEMO_VEHICLES, **DEMO)
    assert np.array_equal(meas.view(np.uint32), g["measurements"].view(np.uint32))
    assert np.array_equal(states.view(np.uint32), g["vehicle_states"].view(np.uint32))
    assert np.array_equal(ego.view(np.uint32), g["ego_pose"].view(np.uint32))
    labels, n_clusters = mine.dbscan(g["dbscan_xy"], 3.0, 5)
    assert n_clusters == int(g["dbscan_clusters"]) and np.array_equal(labels, g["dbscan_labels"])
    steps = [g["eval_cells"][g["eval_offsets"][s] : g["eval_offsets"][s + 1]] for s in range(14)]
    out = evaluate(mine, steps, 250, tools_mod.DEMO_VEHICLES)
    assert out["detections"] == int(g["eval_detections"]) and out["unassigned"] == int(g["eval_unassigned"])
    assert np.array_equal(out["mae"].view(np.uint32), g["eval_mae"].view(np.uint32))
    assert np.array_equal(out["rmse"].view(np.uint32), g["eval_rmse"].view(np.uint32))
