#!/bin/bash
# Quick GPU iteration: tf32 detector parity tests + device-only bench (no CPU legs).  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
for n, k in sorted(d['kernels'].items(), key=lambda x: -x[1]['ms_per_step']):
    print('  %-24s %8.3f ms  %s %s' % (n, k['ms_per_step'], k['bound'], k['frac']))
PY
