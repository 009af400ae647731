#!/bin/bash
# One ncu --set full capture (with source) of the kernels matching $NCU_KERNEL.  Output: gpurun_out/$NCU_OUT.ncu-rep
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-branch_kernel}" -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-1} \
    -f -o gpurun_out/${NCU_OUT:-prof} python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
