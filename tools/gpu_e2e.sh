#!/bin/bash
# e2e variability probe: the same bench under different host-side settings, on ONE box
run() { echo "== $*"; env "$@" DBL_E2E_DIAG=1 python bench.py --no-cpu-baseline 2>&1 >/tmp/line.json | grep "e2e rep"; python -c "
import json; d=json.loads(open('/tmp/line.json').read().strip().splitlines()[-1]); print('   value', d['value'], 'e2e', d['e2e']['value'], 'fp32', d['e2e']['fp32_input']['value'])"; }
run A=1
run DBL_NO_CLK=1
run DBL_E2E_LAG=2
run DBL_NO_CLK=1 DBL_E2E_LAG=2
run A=1
nproc; cat /proc/loadavg; python - <<'PY'
import os; print('affinity', len(os.sched_getaffinity(0)))
try: print(open('/sys/fs/cgroup/cpu.max').read())
except Exception as e: print(e)
PY
