#!/bin/bash
# Full GPU pass (runs ON the GPU box under gpurun): GPU suite, bench lines (N=1, reference arm, configs 2/4/5), kernel tables.
#   bash tools/gpu_full.sh <tag> [quick]
set -u
R=${1:-r02}
O=gpurun_out/$R
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -4 $O/gpu_tests.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline --steps 5 --kernel-table $O/kernel_table.json > /dev/null 2> $O/ktable.err
if [ "${2:-}" != "quick" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
  timeout 300 python bench.py --workload decode_nms --regime typical > $O/bench_config2_typical.json 2> $O/c2t.err
  timeout 300 python bench.py --workload decode_nms --regime dense --steps 10 > $O/bench_config2_dense.json 2> $O/c2d.err
  timeout 600 python bench.py --scale l --batch 32 --steps 10 --no-cpu-baseline --kernel-table $O/kernel_table_l.json > $O/bench_config4_l_b32.json 2> $O/c4.err
  timeout 600 python bench.py --imgsz 1280 --batch 16 --conf 0.001 --steps 10 --no-cpu-baseline > $O/bench_config5_1280_b16.json 2> $O/c5.err
fi
python - <<PY
import json
d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'lat', d['latency_single_step_ms'], 'parity', d.get('parity'))
for k in d['kernels']: print('  ', k)
PY
