"""Host arithmetic of the band mode (no GPU): the slot ranges that consecutive bands derive from the shared prefix sums tile
the global ranges exactly - birth slots [0, int(total * scale)) and output slots [0, N) - for random splits of the mass, for
systematic and stratified offsets, and when the prefixes are exchanged between two processes (gloo, world size 2)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

from _loader import ROOT, load_dogm_b200


def ranges(gpu, lib, locals_, n_glob, seed, cycle, mode):
    before, acc = [], 0.0
    for v in locals_:
        before.append(acc)
        acc = acc + float(v)
    out = []
    for b, v in zip(before, locals_):
        lo, hi = C.c_longlong(0), C.c_longlong(0)
        assert lib.dogm_band_output_range(seed, cycle, mode, n_glob, b, float(v), acc, C.byref(lo), C.byref(hi)) == 0
        out.append((lo.value, hi.value))
    return out, before, acc


def test_output_ranges_tile_the_global_slots():
    gpu = load_dogm_b200()
    lib = gpu.load_library()
    rng = np.random.default_rng(0)
    for case in range(200):
        R = int(rng.integers(1, 9))
        n_glob = int(rng.choice([1, 7, 1000, 2_000_000, 200_000_000]))
        w = rng.uniform(0.0, 1.0, R) ** 4 * rng.choice([1e-3, 1.0, 1e4])
        if case % 5 == 0:
            w[rng.integers(0, R)] = 0.0  # an empty band
        if w.sum() <= 0:
            w[0] = 1.0
        for mode in (gpu.RESAMPLE_SYSTEMATIC, gpu.RESAMPLE_STRATIFIED if n_glob <= 2_000_000 else gpu.RESAMPLE_SYSTEMATIC):
            rs, before, total = ranges(gpu, lib, w, n_glob, 123456 + case, case, mode)
            assert rs[0][0] == 0
            nonzero = [r for r, v in zip(rs, w) if v > 0]
            for (a0, a1), (b0, b1) in zip(nonzero[:-1], nonzero[1:]):
                assert a1 == b0, (case, rs)
            assert nonzero[-1][1] == n_glob
            for (lo, hi), v in zip(rs, w):
                assert lo <= hi and (v > 0 or lo == hi)
                # a band's share of the slots follows its share of the weight (systematic: within one slot)
                assert abs((hi - lo) - v / total * n_glob) <= 1.0 + 1e-9 * n_glob or mode != gpu.RESAMPLE_SYSTEMATIC


def test_birth_slot_ranges_tile():
    gpu = load_dogm_b200()
    lib = gpu.load_library()
    rng = np.random.default_rng(1)
    for case in range(200):
        R = int(rng.integers(1, 9))
        m = rng.uniform(0.0, 50.0, R)
        b_glob = int(rng.choice([0, 1, 333, 200_000, 20_000_000]))
        before, acc = [], 0.0
        for v in m:
            before.append(acc)
            acc = acc + float(v)
        prev_past = 0
        for r in range(R):
            f, p = C.c_int(0), C.c_int(0)
            assert lib.dogm_band_slot_range(before[r], float(m[r]), acc, b_glob, C.byref(f), C.byref(p)) == 0
            assert f.value == prev_past and p.value >= f.value
            prev_past = p.value
        assert b_glob - 1 <= prev_past <= b_glob or acc <= 0  # int(float(total) * (B / total)) is B or B - 1


WORKER = r"""
import ctypes as C, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from _loader import load_dogm_b200
gpu = load_dogm_b200(); lib = gpu.load_library()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
local = torch.tensor([3.25 if rank == 0 else 7.5], dtype=torch.float64)
parts = [torch.zeros(1, dtype=torch.float64) for _ in range(2)]
dist.all_gather(parts, local)                     # the one exchange of the resampling normaliser: a double per band
before, acc = [], 0.0
for p in parts:
    before.append(acc); acc = acc + float(p[0])
lo, hi = C.c_longlong(0), C.c_longlong(0)
lib.dogm_band_output_range(123456, 5, gpu.RESAMPLE_SYSTEMATIC, 1000003, before[rank], float(local[0]), acc, C.byref(lo), C.byref(hi))
mine = torch.tensor([lo.value, hi.value], dtype=torch.int64)
both = [torch.zeros(2, dtype=torch.int64) for _ in range(2)]
dist.all_gather(both, mine)
if rank == 0:
    assert both[0][0] == 0 and both[0][1] == both[1][0] and both[1][1] == 1000003, both
    print("tiling ok", [b.tolist() for b in both])
dist.destroy_process_group()
"""


def test_two_processes_agree_on_the_partition(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29600 + os.getpid() % 200)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "tiling ok" in outs[0]


def test_band_edges_by_phase():
    """balanced_rows_by_phase: the split tiles the rows, keeps the minimum band height, and its modelled cycle time (slowest band
    per phase, summed over the phases) is never worse than that of a split by one load figure per row - clearly better for a
    scene whose particles sit in a part of the grid (a band of empty rows is cheap in particles, expensive in cells)."""
    gpu = load_dogm_b200()
    G = 4096
    y = np.arange(G, dtype=np.float64)
    fan = np.exp(-(G - y) / 700.0)
    fan[:1400] *= 0.02  # far rows: hardly any particles
    fan = fan / fan.sum() * 2e7
    uniform = np.full(G, 2e7 / G)
    for hist in (fan, uniform):
        for R in (2, 3, 4, 8):
            rows = gpu.balanced_rows_by_phase(hist, R)
            assert len(rows) == R and sum(rows) == G and min(rows) >= 64
            load = hist / hist.sum() * (240.0 * 2e7 + 37.0 * 2e6) + 100.0 * G
            ref = gpu.balanced_rows(load, R)
            t_new, t_ref = gpu.band_cycle_model(hist, rows), gpu.band_cycle_model(hist, ref)
            assert t_new <= t_ref * 1.001, (R, t_new, t_ref)
    rows8 = gpu.balanced_rows_by_phase(fan, 8)
    assert gpu.band_cycle_model(fan, rows8) < 0.95 * gpu.band_cycle_model(fan, gpu.balanced_rows(fan / fan.sum() * (240.0 * 2e7 + 37.0 * 2e6) + 100.0 * G, 8))
    # degenerate: no particles at all -> still a valid split
    rows0 = gpu.balanced_rows_by_phase(np.zeros(G), 4)
    assert len(rows0) == 4 and sum(rows0) == G and min(rows0) >= 64


def test_band_edges_from_measured_costs():
    """Closing the loop on the band edges: a synthetic 'machine' whose cell kernel is five times cheaper on rows without
    particles (the quiet-block shortcut) times a placement; band_costs_from_measurement turns the bands' own stage times into
    per-particle costs and one cell cost per band; the placement made with them is faster on that machine than the one made
    with the default costs, and the per-row cell costs are accepted by the solver (tiles the rows, minimum height kept)."""
    gpu = load_dogm_b200()
    G, R = 4096, 8
    y = np.arange(G, dtype=np.float64)
    hist = np.exp(-(G - y) / 500.0)
    hist[:2000] = 0.0
    hist = hist / hist.sum() * 2e7

    def machine(rows):
        """per band [predict, waits, update, birth + CDF, resample] in ms and the cycle time (device-paced: max over the bands of
        predict + update, then of birth + CDF, then of resampling)"""
        cuts = np.concatenate([[0], np.cumsum(rows)]).astype(int)
        ms = []
        for b in range(len(rows)):
            n = hist[cuts[b]:cuts[b + 1]].sum()
            quiet = (hist[cuts[b]:cuts[b + 1]] == 0).sum() * G
            busy = rows[b] * G - quiet
            ms.append([12e-9 * n, 0.0, 37e-9 * n + 25e-9 * busy + 5e-9 * quiet, 13e-9 * n, 17e-9 * n])
        ms = np.asarray(ms)
        return ms, (ms[:, 0] + ms[:, 2]).max() + ms[:, 3].max() + ms[:, 4].max()

    rows1 = gpu.balanced_rows_by_phase(hist, R, **gpu.DEVICE_PACED_COSTS)
    ms1, t1 = machine(rows1)
    cuts = np.concatenate([[0], np.cumsum(rows1)]).astype(int)
    n_b = [hist[cuts[b]:cuts[b + 1]].sum() for b in range(R)]
    costs = gpu.band_costs_from_measurement(n_b, rows1, ms1, G)
    assert costs is not None and costs[2].shape == (G,)
    assert abs(costs[0] - 30.0) < 1.0 and abs(costs[1] - 49.0) < 3.0
    rows2 = gpu.balanced_rows_by_phase(hist, R, cost_particle_phases=costs[0], cost_particle_update=costs[1], cost_cell_update=costs[2])
    assert len(rows2) == R and sum(rows2) == G and min(rows2) >= 64
    _, t2 = machine(rows2)
    assert t2 < 0.97 * t1, (t1, t2, rows1, rows2)
    assert gpu.band_costs_from_measurement([0.0] * R, rows1, ms1, G) is None


def test_two_bands_per_gpu_plan():
    """paired_band_plan: 2 R contiguous bands, band b and band 2 R - 1 - b on one GPU; the split tiles the rows with the minimum
    height, every GPU gets two bands, and for a scene that is dense at one end the modelled cycle (slowest GPU per phase, summed)
    beats the best split with one contiguous band per GPU."""
    gpu = load_dogm_b200()
    G, R = 16384, 8
    y = np.arange(G, dtype=np.float64)
    hist = np.exp(-(G - y) / 2500.0)
    hist[:3000] *= 0.01
    hist = hist / hist.sum() * 2e8
    rows, devices, t_pair = gpu.paired_band_plan(hist, R)
    assert len(rows) == 2 * R and sum(rows) == G and min(rows) >= 64
    assert sorted(devices) == sorted(list(range(R)) * 2) and devices[0] == devices[-1] and devices[R - 1] == devices[R]
    single = gpu.balanced_rows_by_phase(hist, R, **gpu.DEVICE_PACED_COSTS)
    t_single = gpu.band_cycle_model(hist, single, **gpu.DEVICE_PACED_COSTS)
    assert t_pair < 0.93 * t_single, (t_pair, t_single)
    # uniform scene: nothing to gain, nothing lost
    uni = np.full(G, 2e8 / G)
    _, _, t_u = gpu.paired_band_plan(uni, R)
    assert t_u <= 1.01 * gpu.band_cycle_model(uni, gpu.balanced_rows_by_phase(uni, R, **gpu.DEVICE_PACED_COSTS), **gpu.DEVICE_PACED_COSTS)
