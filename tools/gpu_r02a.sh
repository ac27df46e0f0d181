#!/bin/bash
# Round-2 GPU pass A (runs ON the GPU box under gpurun): full GPU suite with measured tolerances printed, bench lines for configs
# 2-5, reference arm, kernel table, compute-sanitizer memcheck + racecheck on the small-shape op tests.
set -u
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -5 $O/gpu_tests.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
timeout 300 python bench.py --no-cpu-baseline --steps 5 --kernel-table $O/kernel_table.json > /dev/null 2> $O/ktable.err
timeout 300 python bench.py --workload decode_nms --regime typical > $O/bench_config2_typical.json 2> $O/c2t.err
timeout 300 python bench.py --workload decode_nms --regime dense --steps 10 > $O/bench_config2_dense.json 2> $O/c2d.err
timeout 300 python bench.py --workload decode_nms --regime typical --nc 3 > $O/bench_config2_typical_nc3.json 2> $O/c2t3.err
timeout 600 python bench.py --scale l --batch 32 --steps 10 --no-cpu-baseline --kernel-table $O/kernel_table_l.json > $O/bench_config4_l_b32.json 2> $O/c4.err
timeout 600 python bench.py --imgsz 1280 --batch 16 --conf 0.001 --steps 10 --no-cpu-baseline > $O/bench_config5_1280_b16.json 2> $O/c5.err
# compute-sanitizer: the per-kernel op tests + the small whole-model tests, memcheck then racecheck
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nms.py -m gpu -q -x -p no:cacheprovider > $O/sanitizer_memcheck_ops.log 2>&1; echo "memcheck ops rc=$?" | tee -a $O/sanitizer_memcheck_ops.log
timeout 600 $CS --tool memcheck --error-exitcode 99 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?" | tee -a $O/sanitizer_memcheck_smoke.log
timeout 900 $CS --tool racecheck --error-exitcode 99 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?" | tee -a $O/sanitizer_racecheck_smoke.log
ls -la $O
