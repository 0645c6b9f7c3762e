#!/bin/bash
# End-of-round evidence on one B200: ncu launch list of a bench run, ncu full set of two cycles (headline and highway
# configurations), launch list of the reference arm, free-running timeline.  Outputs under gpurun_out/ (turned into the committed
# summaries under profiles/ by tools/ncu_summary.py and tools/ref_launch_sum.py, which run without a GPU).
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-band-scaling > gpurun_out/bench_under_ncu_$TAG.log 2>&1
SKIP=$(python tools/profile_cycle.py --print-skip 2>/dev/null | tail -1)
ncu --set full --clock-control none --import-source on --launch-skip $SKIP -c 30 -f -o gpurun_out/prof_${TAG}_nuss \
    python tools/profile_cycle.py > gpurun_out/ncu_full_$TAG.log 2>&1
SKIPH=$(python tools/profile_cycle.py --config highway --warm 6 --print-skip 2>/dev/null | tail -1)
ncu --set full --clock-control none --launch-skip $SKIPH -c 30 -f -o gpurun_out/prof_${TAG}_highway \
    python tools/profile_cycle.py --config highway --warm 6 > gpurun_out/ncu_full_${TAG}_highway.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ref_launches_$TAG.csv \
    python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/bench_ref_under_ncu_$TAG.log 2>&1
python tools/trace_cycle.py nuss 9 > gpurun_out/trace_$TAG.txt 2>&1
# summaries on the box (ncu -i needs no GPU, but the reports are too large to travel back together)
python tools/ncu_summary.py gpurun_out/prof_${TAG}_nuss.ncu-rep gpurun_out/${TAG} nuss 2 > gpurun_out/summary_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_${TAG}_highway.ncu-rep gpurun_out/${TAG}_highway highway 2 >> gpurun_out/summary_$TAG.log 2>&1
python tools/ref_launch_sum.py gpurun_out/ref_launches_$TAG.csv nuss >> gpurun_out/summary_$TAG.log 2>&1
cp profiles/ref_kernel_sum.json profiles/r02_ref_cycle_nuss.csv gpurun_out/ 2>/dev/null
rm -f gpurun_out/prof_${TAG}_highway.ncu-rep
tail -3 gpurun_out/ncu_full_$TAG.log; tail -3 gpurun_out/ncu_full_${TAG}_highway.log; tail -2 gpurun_out/bench_ref_under_ncu_$TAG.log; cat gpurun_out/trace_$TAG.txt; ls -la gpurun_out/*$TAG*
