#!/bin/bash
# ncu full captures of the secondary rows: HardNet tensor-core kernels (4096 patches) and the SMNN tcgen05 kernel (2048^2).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"hn_tc_" -s 9 -c 9 -f -o gpurun_out/prof_hn python scripts/hn_bench.py > gpurun_out/ncu_hn.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"nn2_tc_kernel|merge_splits_kernel|smnn_select_kernel" -s 4 -c 4 -f -o gpurun_out/prof_smnn python scripts/smnn_bench.py > gpurun_out/ncu_smnn.log 2>&1
for r in prof_hn prof_smnn; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out | tail -8
