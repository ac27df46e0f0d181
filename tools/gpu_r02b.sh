#!/bin/bash
# GPU pass B: the corrected parity tests + e2e pipeline diagnosis
set -u
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_model.py -m gpu -q -rA -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
for lag in 0 1 2; do
  DBL_E2E_LAG=$lag timeout 300 python bench.py --no-cpu-baseline > $O/bench_lag$lag.json 2> $O/bench_lag$lag.err
done
python - <<'PY' > $O/h2d.txt 2>&1
import torch, time
x = torch.empty(64,3,640,640, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device='cuda')
for n in range(3):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(x, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print('H2D u8 78.6MB: %.3f ms -> %.1f GB/s' % (e0.elapsed_time(e1)/10, x.numel()/ (e0.elapsed_time(e1)/10*1e-3)/1e9))
PY
cat $O/h2d.txt
python -c "
import json
for l in (0,1,2):
    d=json.loads(open('$O/bench_lag%d.json'%l).read().strip().splitlines()[-1]); print(l, d['value'], d['e2e']['value'], d['e2e']['fp32_input']['value'], d['latency_single_step_ms'])
"
