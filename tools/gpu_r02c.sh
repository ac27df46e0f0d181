#!/bin/bash
set -u
O=gpurun_out/r02c
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "dsconv_tc" > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -8 $O/gpu_tests.log
timeout 300 python tools/ds_bench.py > $O/ds_bench.txt 2>&1; tail -30 $O/ds_bench.txt
