"""Sum of the GPU kernel durations of one reference updateGrid cycle from an ncu launch list of the reference arm:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ref_launches.csv \
        python bench.py --impl reference --steps 4 --warmup 3 [--config nuss]
    python tools/ref_launch_sum.py gpurun_out/ref_launches.csv nuss        -> profiles/ref_kernel_sum.json, profiles/r02_ref_cycle.csv
A cycle is the stretch between two launches of the reference's predictKernel (one per updateGrid); the last complete cycles
are averaged.  ncu serialises the launches and replays them cold, so the sum is an upper bound of the kernels' share."""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, config = sys.argv[1], sys.argv[2]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns").lower()
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], us))
    marks = [i for i, (k, _) in enumerate(rows) if "predictKernel" in k]
    if len(marks) < 3:
        raise SystemExit(f"only {len(marks)} cycles in {path}")
    cycles = [(a, b) for a, b in zip(marks[:-1], marks[1:])][-3:]
    sums = [sum(us for _, us in rows[a:b]) for a, b in cycles]
    per_kernel = OrderedDict()
    a, b = cycles[-1]
    for k, us in rows[a:b]:
        short = k.split("(")[0].split("<")[0].strip()
        e = per_kernel.setdefault(short, [0, 0.0])
        e[0] += 1
        e[1] += us
    out = os.path.join(ROOT, "profiles", "ref_kernel_sum.json")
    data = {}
    if os.path.exists(out):
        data = json.load(open(out))
    data[config] = sum(sums) / len(sums) * 1e-3
    json.dump(data, open(out, "w"), indent=1, sort_keys=True)
    with open(os.path.join(ROOT, "profiles", f"r02_ref_cycle_{config}.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches_per_cycle", "total_us"])
        for k, (n, us) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{us:.2f}"])
        w.writerow(["SUM", b - a, f"{sums[-1]:.2f}"])
    print(config, "kernel sum per cycle [ms]:", [round(s * 1e-3, 3) for s in sums], "launches per cycle:", b - a)


if __name__ == "__main__":
    main()
