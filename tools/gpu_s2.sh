#!/bin/bash
# Runs ON THE GPU BOX: parity of the thin stride-2 tcgen05 kernel, its micro-benchmark, then the model-level check.
O=gpurun_out/s2
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "stride2_pixel_pair" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/tests.log
timeout 120 python tools/conv_bench.py --impl 0 --only s2_ > $O/conv_bench.txt 2>&1; cat $O/conv_bench.txt
DBL_S2_TC=0 timeout 120 python tools/conv_bench.py --impl 0 --only s2_ 2>&1 | sed 's/^/old: /'
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -q -x > $O/model.log 2>&1; echo "model rc=$?"; tail -3 $O/model.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --kernel-table $O/kt.json > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2/bench.json').read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
t=json.load(open('gpurun_out/s2/kt.json'))
for r in t['launches']:
    if r['what'] in ('model.3','model.5','model.12.fuse.down','model.9.cv2'): print(r['what'], r['kern'], round(r['ms']*1e3,1))
PY
