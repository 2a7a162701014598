#!/bin/bash
# compute-sanitizer over the rasterizer kernels on small cases (memcheck / racecheck / synccheck).  bash tools/gpu_sanitize.sh [tag]
tag=${1:-san}; out=gpurun_out/$tag; mkdir -p $out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=ggrt python tools/sanitize_case.py > $out/$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/$tool.log | tail -1)"
  grep -E "^case" $out/$tool.log | head -4
done
