#!/bin/bash
# Runs tools/sanitize_target.py under the four compute-sanitizer tools on the GPU box and leaves the logs in
# gpurun_out/sanitizer_<tool>.log. Usage (from the repo root): gpurun -- 'bash tools/sanitize.sh'
mkdir -p gpurun_out
rc=0
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  code=$?
  echo "$tool exit $code: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
  [ $code -ne 0 ] && rc=$code
done
exit $rc
