#!/bin/bash
# A/B of library builds on the headline config: tools/ab.sh <name|-> ...   ("-" = the in-tree library, name = build_ab/lib_<name>.so)
for v in "$@"; do
  if [ "$v" = "-" ]; then unset DOGM_B200_LIB; else export DOGM_B200_LIB=build_ab/lib_$v.so; fi
  timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ab_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("$v", round(d["value"],1), round(d["ms_per_step"]*1000,1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["cycle"]["frac"],3))
print("   ", {k:round(v*1000,1) for k,v in d["roofline"]["kernels_ms_per_cycle"].items()})
PY
done
