mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms15_(tma_)?kernel" -s 3 -c 1 -f -o gpurun_out/nms15 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_nms15.log 2>&1
ls -la gpurun_out/nms15.ncu-rep
