#!/bin/bash
# Late round-2 measurement set (one gpurun call): bench of record + reference arm, ncu launch list, full capture of the
# windowed NMS + select kernels (the only kernels changed since r02g; raw pages exported).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_raw.csv gpurun_out/prof_greedy_raw.csv gpurun_out/prof_hn_raw.csv gpurun_out/prof_smnn_raw.csv gpurun_out/prof_x3_raw.csv
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"nms15_tma_kernel|nms15_kernel|select_sort_kernel" -s 6 -c 2 -f -o gpurun_out/prof_nms $B > gpurun_out/ncu_nms.log 2>&1
ncu -i gpurun_out/prof_nms.ncu-rep --page raw --csv > gpurun_out/prof_nms_raw.csv 2>/dev/null
rm -f gpurun_out/prof_nms.ncu-rep
tail -c 600 gpurun_out/bench.json; ls -la gpurun_out | tail -12
