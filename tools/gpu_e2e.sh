#!/bin/bash
# e2e variability probe: many repetitions in ONE process, with the GPU's PCIe link / clocks / power sampled beside them
(while true; do nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv,noheader; sleep 0.25; done) > /tmp/smi.log 2>&1 &
SMI=$!
for i in 1 2 3; do
  DBL_E2E_DIAG=1 DBL_E2E_REPS=15 python bench.py --no-cpu-baseline 2>&1 >/tmp/line.json | grep "e2e rep"
  python -c "
import json; d=json.loads(open('/tmp/line.json').read().strip().splitlines()[-1]); print('   value', d['value'], 'e2e', d['e2e']['value'], 'fp32', d['e2e']['fp32_input']['value'])"
done
kill $SMI
sort /tmp/smi.log | uniq -c | sort -rn | head -12
