#!/bin/bash
# Source-level ncu captures (per-line counters incl. shared-memory bank conflicts) of the stage-1 detector kernels.
# Only the exported source pages come back (the reports are too large).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_branch_kernel" -s 8 -c 2 -f -o gpurun_out/src_branch $B > gpurun_out/ncu_src.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_merge_bulk_kernel" -s 2 -c 1 -f -o gpurun_out/src_merge $B >> gpurun_out/ncu_src.log 2>&1
for r in src_branch src_merge; do
  ncu -i gpurun_out/$r.ncu-rep --page source --print-source sass,cuda --csv > gpurun_out/${r}_cuda.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out/src_* | cut -c1-120
