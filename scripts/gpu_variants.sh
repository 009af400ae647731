#!/bin/bash
# A/B development helper: device-only bench per library variant under build/variants/.
for v in "$@"; do
  echo "== $v"
  BALF_B200_LIB=$PWD/build/variants/$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'],1))
for n,k in sorted(d['kernels'].items(), key=lambda x:-x[1]['ms_per_step'])[:${TOPN:-8}]:
    print('  %-24s %8.3f ms' % (n, k['ms_per_step']))
"
done
