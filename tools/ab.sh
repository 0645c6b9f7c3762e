#!/bin/bash
# A/B of library builds on the headline config: tools/ab.sh <name|-> ...   ("-" = the in-tree library, name = build_ab/lib_<name>.so)
STEPS=${AB_STEPS:-100}
for v in "$@"; do
  if [ "$v" = "-" ]; then unset DOGM_B200_LIB; else export DOGM_B200_LIB=build_ab/lib_$v.so; fi
  timeout 300 python bench.py --steps $STEPS --warmup 10 --no-cpu-baseline --no-band-scaling ${AB_ARGS} 2>gpurun_out/ab_$v.err | tail -1 > gpurun_out/ab_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("$v", round(d["value"],1), "cycles/s", round(d["ms_per_step"]*1000,1), "us; e2e", round(d["e2e"]["value"],1), "stationary", round(d["protocol"]["stationary_cycle_ms"]*1000,1), "launches/cycle", d["gpu_launches"]/d["steps"])
print("   timeline", {k:round(v,1) for k,v in d["roofline"]["timeline_us"].items()})
print("   events  ", {k:round(v*1000,1) for k,v in d["roofline"]["kernels_ms_per_cycle"].items()})
PY
done
