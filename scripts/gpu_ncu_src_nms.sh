set -u
mkdir -p gpurun_out
B="python scripts/nms_bench.py 64"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms15_tma_kernel|nms15_kernel|select_sort_kernel" -s 6 -c 2 -f -o gpurun_out/src_nms $B > gpurun_out/ncu_src.log 2>&1
ncu -i gpurun_out/src_nms.ncu-rep --page source --print-source sass,cuda --csv > gpurun_out/src_nms_cuda.csv 2>/dev/null
ncu -i gpurun_out/src_nms.ncu-rep --page raw --csv > gpurun_out/src_nms_raw.csv 2>/dev/null
rm -f gpurun_out/src_nms.ncu-rep
ls -la gpurun_out/src_nms* | cut -c1-120
