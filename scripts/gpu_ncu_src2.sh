set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_branch_kernel" -s 10 -c 2 -f -o gpurun_out/src_b64 $B > gpurun_out/ncu_src.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_merge_bulk_kernel" -s 3 -c 1 -f -o gpurun_out/src_m64 $B >> gpurun_out/ncu_src.log 2>&1
for r in src_b64 src_m64; do
  ncu -i gpurun_out/$r.ncu-rep --page source --print-source sass,cuda --csv > gpurun_out/${r}_cuda.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out/src_* | cut -c1-120
