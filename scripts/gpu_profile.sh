#!/bin/bash
# ncu evidence for profiles/: launch list of the bench of record, full captures of the dominant detector kernel,
# the NMS kernel and the HardNet conv kernels.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_merge_kernel|tc_branch_kernel" -s 12 -c 3 \
    -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms15_(tma_)?kernel|select_sort" -s 4 -c 2 \
    -f -o gpurun_out/prof_nms python scripts/nms_bench.py 64 > gpurun_out/ncu_nms.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hn_tc" -s 16 -c 8 \
    -f -o gpurun_out/prof_hn python scripts/hn_bench.py 4096 tf32 > gpurun_out/ncu_hn.log 2>&1
ls -la gpurun_out
