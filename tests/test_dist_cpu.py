"""World-size-2 test (gloo, CPU) of bench.py's multi-GPU plumbing: one process per rank, independent sensor streams
per rank (replicas: no data-path collective), barrier + max-over-ranks timing, whole-job aggregate on rank 0."""
import json
import os
import socket
import subprocess
import sys
import textwrap

from _loader import ROOT

WORKER = textwrap.dedent(
    """
    import json, os, sys, time
    sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
    import numpy as np
    import bench
    from _loader import load_oracle
    d = bench.Dist(2)
    orc = load_oracle()
    cfg = bench.CONFIGS["tiny"]
    # every rank owns one sensor stream: its own scene seed, its own filter instance
    beams = bench.make_beams(cfg, 6, seed=1234 + d.rank)
    laser = orc.LaserParams(cfg["size"], cfg["resolution"], bench.FOV, bench.STDDEV_RANGE)
    p = orc.Params(cfg["size"], cfg["resolution"], cfg["n"], cfg["b"], *bench.DEMO_PARAMS)
    o = orc.OracleDOGM(p, resample_mode=orc.RESAMPLE_SYSTEMATIC)
    rng = np.random.default_rng(d.rank)
    n, b = cfg["n"], cfg["b"]
    d.barrier()
    t0 = time.perf_counter()
    K = 4
    for c in range(K):
        meas = orc.meas_generate(laser, cfg["size"], cfg["resolution"], beams[c])
        o.set_noise(rng.standard_normal((n, 4)).astype(np.float32), rng.standard_normal((b, 2)).astype(np.float32),
                    rng.uniform(-1, 1, (n, 2)).astype(np.float32), np.array([0.5], np.float32))
        x, y = bench.pose_at(c)
        o.update_grid(meas, float(x), float(y), 0.0, bench.DT)
    t = time.perf_counter() - t0
    t_max = d.max(t)
    occ = float(o.grid_cells["occ_mass"].sum())
    occ_sum = d.sum(occ)
    assert t_max >= t - 1e-9
    if d.rank == 0:
        print(json.dumps({"world": d.world, "value": d.world * K / t_max, "t_max": t_max, "occ_sum": occ_sum, "occ0": occ,
                          "beams0": float(np.nan_to_num(beams[0], posinf=0).sum())}))
    else:
        print(json.dumps({"rank": d.rank, "occ": occ, "beams0": float(np.nan_to_num(beams[0], posinf=0).sum())}))
    d.close()
    """
)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_replicas_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), DOGM_BENCH_DIST_BACKEND="gloo", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    r0 = json.loads(outs[0][0].strip().splitlines()[-1])
    r1 = json.loads(outs[1][0].strip().splitlines()[-1])
    assert r0["world"] == 2 and r0["value"] > 0
    # the two ranks ran different streams (different scene seeds) and the aggregate saw both
    assert r0["beams0"] != r1["beams0"]
    assert abs(r0["occ_sum"] - (r0["occ0"] + r1["occ"])) < 1e-3 * max(1.0, r0["occ_sum"])


def test_reference_arm_runs_on_rank_zero_only():
    """bench.py --impl reference under a 2-rank launch: rank 1 exits without work and prints nothing."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()))
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "3", "--config", "tiny"], env=env, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""
