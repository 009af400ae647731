#!/bin/bash
# Round-2 measurement set (one gpurun call): bench of record + reference arm, ncu launch list, full captures of the detector,
# windowed NMS, greedy NMS, HardNet and SMNN kernels (raw pages exported; the .ncu-rep files are too large to bring back).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"tc_branch_kernel|tc_merge_bulk_kernel|tc_merge_kernel|tc_head_kernel|pool_kernel" -s 48 -c 16 \
    -f -o gpurun_out/prof $B > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"nms15_(tma_)?kernel|select_sort_kernel" -s 6 -c 2 -f -o gpurun_out/prof_nms $B > gpurun_out/ncu_nms.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"greedy_cells" -s 24 -c 8 -f -o gpurun_out/prof_greedy $B --nms greedy --precision tf32 > gpurun_out/ncu_greedy.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"hn_tc_(conv|first|final)_kernel" -s 7 -c 7 -f -o gpurun_out/prof_hn python scripts/hn_bench.py > gpurun_out/ncu_hn.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"nn2_tc_kernel|merge_splits_kernel|smnn_select_kernel" -s 4 -c 4 -f -o gpurun_out/prof_smnn python scripts/smnn_bench.py > gpurun_out/ncu_smnn.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"tc_branch_kernel|tc_merge_bulk_kernel|tc_merge_kernel|tc_head_kernel" -s 45 -c 15 \
    -f -o gpurun_out/prof_x3 $B --nms greedy > gpurun_out/ncu_x3.log 2>&1
for r in prof prof_nms prof_greedy prof_hn prof_smnn prof_x3; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out | tail -20
