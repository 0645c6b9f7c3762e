"""Records tests/golden/demo_tools.npz from the reference's own simulator / DBSCAN / evaluator classes
(oracle/_ref/libdemo_ref.so, built by oracle/build_demo_ref.sh from /root/reference).  Run from the repo root."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_tools_cpu as T  # noqa: E402

ref = T.tools_mod.Tools(T.REF_LIB, "ref_tools_")
meas, states, ego = ref.simulate(vehicles=T.tools_mod.DEMO_VEHICLES, **T.DEMO)
rng = np.random.default_rng(11)
steps = T.demo_like_cells(rng, ref)
out = T.evaluate(ref, steps, 250, T.tools_mod.DEMO_VEHICLES)
cells0 = np.sort(steps[3], order="cell_idx")
xy = np.stack([cells0["cell_idx"] % 250, cells0["cell_idx"] // 250], axis=1).astype(np.float32)
labels, n_clusters = ref.dbscan(xy, 3.0, 5)
offsets = np.concatenate([[0], np.cumsum([len(s) for s in steps])]).astype(np.int64)
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "demo_tools.npz"),
    measurements=meas, vehicle_states=states, ego_pose=ego, dbscan_xy=xy, dbscan_labels=labels, dbscan_clusters=n_clusters,
    eval_cells=np.concatenate(steps), eval_offsets=offsets, eval_mae=out["mae"], eval_rmse=out["rmse"],
    eval_detections=out["detections"], eval_unassigned=out["unassigned"],
)
print("wrote demo_tools.npz:", out)
