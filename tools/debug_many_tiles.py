"""Developer tool: where the CUDA path and the oracle part ways at a given size (first cycle)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _loader import load_dogm_b200, load_oracle
from conftest import make_params, synthetic_meas, cycle_noise
gpu, orc = load_dogm_b200(), load_oracle()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_300_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
size, res = 64.0, 0.5
rng = np.random.default_rng(23)
p, po = make_params(gpu, size, res, n, b), make_params(orc, size, res, n, b)
d = gpu.DOGM(p); d.set_options(noise_mode=gpu.NOISE_INJECTED, resample_mode=gpu.RESAMPLE_INJECTED)
o = orc.OracleDOGM(po, resample_mode=orc.RESAMPLE_INJECTED)
for c in range(2):
    meas = synthetic_meas(gpu.MEAS_CELL_DTYPE, d.grid_size, rng)
    pn, bn, iv, ru = cycle_noise(rng, n, b, p)
    d.set_noise(pn, bn, iv, ru); o.set_noise(pn, bn, iv, ru)
    d.update_grid(meas, 0.3 * c, 0.8 * c, 0.0, 0.1, device=False)
    o.update_grid(meas.view(orc.MEAS_CELL_DTYPE), 0.3 * c, 0.8 * c, 0.0, 0.1)
    ga, oa = d.get_resampled_indices(), o.resampled_idx
    bad = np.flatnonzero(ga != oa)
    print(f"cycle {c}: ancestors differ at {bad.size} of {ga.size}", bad[:10], ga[bad[:10]], oa[bad[:10]])
    gw, ow = d.get_weight_array(), o.weight_array
    wb = np.flatnonzero(gw.view(np.uint32) != ow.view(np.uint32))
    print("   weight_array differs at", wb.size, wb[:10], gw[wb[:5]], ow[wb[:5]])
    gc, oc = d.get_joint_weight_accum(), o.joint_weight_accum
    rel = np.abs(gc - oc) / np.maximum(np.abs(oc), 1e-300)
    print("   cdf max rel", rel.max(), "at", int(rel.argmax()), "total", gc[-1], oc[-1])
    gp, op = d.get_particles(), o.particles
    sb = np.flatnonzero((gp.state.view(np.uint32) != op.state.view(np.uint32)).any(axis=1))
    print("   states differ at", sb.size, sb[:10], "cells differ", int((gp.grid_cell_idx != op.grid_cell_idx).sum()))
    bm = np.flatnonzero(d.get_born_masses().view(np.uint32) != o.born_masses.view(np.uint32))
    print("   born masses differ at", bm.size)
    gb, ob = d.get_birth_particles(), o.birth_particles
    print("   birth idx differ", int((gb.grid_cell_idx != ob.grid_cell_idx).sum()), "birth weights differ", int((gb.weight.view(np.uint32) != ob.weight.view(np.uint32)).sum()))
    if bad.size:
        i = bad[0]
        print("   first bad slot", i, "u", ru[i], "gpu anc", ga[i], "oracle anc", oa[i], "cdf around", gc[min(ga[i], oa[i]) - 1:max(ga[i], oa[i]) + 2], oc[min(ga[i], oa[i]) - 1:max(ga[i], oa[i]) + 2])
        print("   joint_max f32", np.float32(gc[-1]), np.float32(oc[-1]), "draw", np.float32(gc[-1]) * ru[i])
    if bad.size or sb.size:
        break
