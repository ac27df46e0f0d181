#!/bin/bash
O=gpurun_out/st
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "stem_pair_tc" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/tests.log
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -q -x > $O/model.log 2>&1; echo "model rc=$?"; tail -3 $O/model.log
bash tools/gpu_kt.sh 'stem'
