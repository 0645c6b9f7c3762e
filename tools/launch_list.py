"""Turns the raw `ncu --metrics gpu__time_duration.sum --csv` log of a bench run into the two committed files
   <out>_launches.csv       one line per launch: id, kernel, block, grid, duration_ns
   <out>_launch_shares.csv  per kernel: launches, total, average, share of the profiled time
Usage: python tools/launch_list.py gpurun_out/launches_r01.csv profiles/r01 "<the profiled command>" """
import csv
import re
import sys


def main():
    src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, data = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = {}
    with open(out + "_launches.csv", "w", newline="") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv {cmd}\n")
        f.write("# (cold caches, serialised launches: shares, not absolute times; the first launches are creation + first-cycle initialisation;\n")
        f.write("#  bench.py also runs its end-to-end, blocking, stationary, concurrent-stream, instrumented and traced loops)\n")
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "duration_ns"])
        for r in data:
            name = re.sub(r"^void ", "", r[col["Kernel Name"]])
            short = re.match(r"(?:dogm_b200::)?([A-Za-z_0-9]+(?:<[^>]*>)?)", name).group(1)
            ns = float(r[col["Metric Value"]].replace(",", ""))
            w.writerow([r[col["ID"]], short, r[col["Block Size"]], r[col["Grid Size"]], int(ns)])
            a = agg.setdefault(short, [0, 0.0])
            a[0] += 1
            a[1] += ns * 1e-3
    total = sum(a[1] for a in agg.values())
    with open(out + "_launch_shares.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], f"{a[1]:.1f}", f"{a[1] / a[0]:.2f}", f"{a[1] / total:.3f}"])
    print("wrote", out + "_launches.csv", out + "_launch_shares.csv")


if __name__ == "__main__":
    main()
