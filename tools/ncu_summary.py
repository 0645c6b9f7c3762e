"""Turns an `ncu --set full` report into the small, committed summaries under profiles/:
   <out>_kernels.csv  one line per profiled launch: duration, DRAM bytes, throughput %, occupancy, IPC, top stalls
   traffic.json       {config: {kernel: {dram_bytes per launch, dram_pct, duration_us, launches_per_cycle}}} (read by bench.py ->
                      roofline.traffic / cycle_dram / per_kernel_dram_pct; other configurations' entries are kept)
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02 [config=nuss] [cycles profiled=2]"""
import csv
import io
import json
import os
import re
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    config = sys.argv[3] if len(sys.argv) > 3 else "nuss"
    cycles = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, name):
        try:
            return float(r[col[name]].replace(",", ""))
        except Exception:
            return float("nan")

    def to_bytes(r, name):
        v, u = num(r, name), units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    def to_us(r, name):
        v, u = num(r, name), units[col[name]].lower()
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)

    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    lines, traffic, agg = [], {}, {}
    for r in data:
        name = re.sub(r"^void ", "", r[col["Kernel Name"]])
        short = re.match(r"(?:dogm_b200::)?([A-Za-z_0-9]+)", name).group(1)
        stalls = sorted(((num(r, c), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stall_cols), reverse=True)[:3]
        dur = to_us(r, "gpu__time_duration.sum")
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        lines.append([
            short, f"{dur:.2f}", f"{rd / 1e6:.2f}", f"{wr / 1e6:.2f}",
            f"{num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}",
            f"{num(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}",
            f"{num(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}",
            f"{num(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}",
            f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}",
            f"{num(r, 'sm__inst_executed.avg.per_cycle_elapsed'):.2f}",
            f"{int(num(r, 'launch__registers_per_thread'))}",
            f"{int(num(r, 'smsp__inst_executed.sum'))}",
            " ".join(f"{n}={v:.1f}" for v, n in stalls),
        ])
        a = agg.setdefault(short, {"launches": 0, "bytes": 0.0, "us": 0.0, "pct": 0.0})
        a["launches"] += 1
        a["bytes"] += rd + wr
        a["us"] += dur
        a["pct"] += num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')
    for k, a in agg.items():
        traffic[k] = {"dram_bytes": a["bytes"] / a["launches"], "dram_pct": round(a["pct"] / a["launches"], 1),
                      "duration_us": round(a["us"] / a["launches"], 2), "launches_per_cycle": a["launches"] / cycles}
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out + "_kernels.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "duration_us", "dram_read_MB", "dram_write_MB", "dram_pct", "l2_pct", "l1tex_pct", "sm_pct",
                    "warps_active_pct", "ipc", "regs", "warp_insts", "top_stalls(per issue)"])
        w.writerows(lines)
    total = sum(a["us"] for a in agg.values())
    with open(out + "_shares.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_of_profiled_time", "dram_bytes_per_launch"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            w.writerow([k, a["launches"], f"{a['us']:.2f}", f"{a['us'] / total:.3f}", f"{a['bytes'] / a['launches']:.0f}"])
    tpath = os.path.join(os.path.dirname(out) or ".", "traffic.json")
    allcfg = {}
    if os.path.exists(tpath):
        try:
            allcfg = json.load(open(tpath))
            if allcfg and not all(isinstance(v, dict) and all(isinstance(x, dict) for x in v.values()) for v in allcfg.values()):
                allcfg = {}  # the round-1 flat format
        except Exception:
            allcfg = {}
    allcfg[config] = traffic
    json.dump(allcfg, open(tpath, "w"), indent=1, sort_keys=True)
    print("wrote", out + "_kernels.csv", out + "_shares.csv", tpath)


if __name__ == "__main__":
    main()
