#!/bin/bash
# Run on the B200 box through gpurun: GPU parity tests, bench of record, ncu launch list,
# one full ncu capture of the dominant kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi.txt
if [ "${BENCH:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
fi
if [ "${NCU:-1}" = "1" ]; then
# launch list of the same command as the bench of record (batch 64), then one full capture of every detector kernel of one step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"${NCU_KERNEL:-tc_branch_kernel|tc_merge_bulk_kernel|tc_merge_kernel|tc_head_kernel|pool_kernel}" -s ${NCU_SKIP:-48} -c ${NCU_COUNT:-16} \
    -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"nms15_(tma_)?kernel|select_sort_kernel" -s 6 -c 2 \
    -f -o gpurun_out/prof_nms python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_nms.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"greedy_cells" -s 24 -c 8 \
    -f -o gpurun_out/prof_greedy python bench.py --steps 1 --warmup 3 --nms greedy --precision tf32 --no-cpu-baseline --no-extras > gpurun_out/ncu_greedy.log 2>&1
# the reports of 1-3 ms kernels exceed what gpurun copies back: keep their raw pages (what scripts/summarize_profiles.py reads)
for r in prof prof_nms prof_greedy; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out
fi
