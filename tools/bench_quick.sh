timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_q.json; python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["cycle"]["frac"])
print({k:round(v*1000,1) for k,v in d["roofline"]["kernels_ms_per_cycle"].items()})
PY
