#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python scratch/sanitize.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitize_$tool.log | tail -1)"
done
