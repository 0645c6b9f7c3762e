#!/bin/bash
# compute-sanitizer, all four tools, over tools/sanitize_run.py; summaries into gpurun_out/sanitize_<tool>.txt
for tool in memcheck initcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize run done' gpurun_out/sanitize_$tool.txt | tr '\n' ' ')"
done
