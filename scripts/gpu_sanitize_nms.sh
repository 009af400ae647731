#!/bin/bash
# compute-sanitizer on the windowed-NMS tests that run the persistent TMA kernel (and its fallback); outputs in gpurun_out/
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_postproc.py -m gpu -q -x \
      -k "tma_and_fallback or windowed_golden or two_level" > gpurun_out/sanitize_nms_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_nms_$tool.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitize_nms_$tool.log | tail -1)"
done
