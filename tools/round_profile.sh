#!/bin/bash
# End-of-round evidence on one B200: GPU test suite, both bench arms, ncu launch list of a bench run, ncu full set of two cycles.
# Outputs under gpurun_out/ (copy the summaries into profiles/ with tools/ncu_summary.py).
TAG=${1:-r01}
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/bench_ref_$TAG.err | tail -1 > gpurun_out/bench_ref_$TAG.json
python bench.py 2> gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
SKIP=$(python tools/profile_cycle.py --print-skip 2>/dev/null | tail -1)
ncu --set full --clock-control none --import-source on --launch-skip $SKIP -c 30 -f -o gpurun_out/prof_${TAG}_final \
    python tools/profile_cycle.py > gpurun_out/ncu_full_$TAG.log 2>&1
python tools/trace_cycle.py nuss 9 > gpurun_out/trace_$TAG.txt 2>&1
cat gpurun_out/pytest_gpu_$TAG.log; tail -c 600 gpurun_out/bench_$TAG.json; echo; tail -3 gpurun_out/ncu_full_$TAG.log; cat gpurun_out/trace_$TAG.txt
