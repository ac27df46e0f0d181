#!/bin/bash
O=gpurun_out/t5
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "single_chunk or stride2_pixel" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/tests.log
timeout 120 python tools/conv_bench.py --impl 0 --only l6cv; DBL_TC5=0 timeout 120 python tools/conv_bench.py --impl 0 --only l6cv | sed 's/^/tc4: /'
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -q -x > $O/model.log 2>&1; echo "model rc=$?"; tail -3 $O/model.log
bash tools/gpu_kt.sh 'conv_tc5|conv_tc4'
