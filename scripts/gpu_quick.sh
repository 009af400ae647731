#!/bin/bash
# Fast GPU iteration: tensor-core detector parity + a short bench (no CPU baseline leg).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detector_tf32.py tests/test_gpu_detector.py -q --tb=line 2>&1 | tail -8
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("value %.1f img/s  e2e %.1f  ms/step %.2f  det %.2f ms  clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["detector"]["ms_per_step"], d["clocks"]))
for k, v in d["kernels"].items():
    print("  %-24s %8.3f ms %5.1f%%" % (k, v["ms_per_step"], 100 * v["share"]))
PY
tail -3 gpurun_out/bench_quick.err
