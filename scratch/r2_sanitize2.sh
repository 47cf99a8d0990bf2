#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  IMC_LAZY_CLEAN_MIN=0 timeout 170 compute-sanitizer --tool $tool --error-exitcode 3 python scratch/sanitize2.py > gpurun_out/r2g_sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2g_sanitize_$tool.log | tail -1) $(grep -c '^[a-z].* [0-9]' gpurun_out/r2g_sanitize_$tool.log) cases"
done
