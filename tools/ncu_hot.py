"""Hot SASS lines of one kernel in an ncu report: python tools/ncu_hot.py rep.ncu-rep <kernel regex> [nth launch] [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
import re
# the csv holds one block per launch: "Kernel Name" line, header line, rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and row:
        cur["rows"].append(row)
blocks = [b for b in blocks if re.search(rx, b["name"])][::2]  # every launch is listed twice
b = blocks[nth]
hdr, data = b["rows"][0], [r for r in b["rows"][1:] if len(r) > 5]
print(b["name"])
samp = hdr.index("# Samples")
ex = hdr.index("Instructions Executed")
tot = sum(int(r[samp]) for r in data)
print("total samples", tot, "instructions", len(data))
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[1]]
prev = 0
for bb in bars + [len(data)]:
    print(f"  region [{prev},{bb}) samples {sum(int(r[samp]) for r in data[prev:bb])}")
    prev = bb
top = sorted(range(len(data)), key=lambda i: -int(data[i][samp]))[:top_n]
for i in sorted(top):
    r = data[i]
    print(i, r[1].strip()[:80], r[samp], r[ex])
