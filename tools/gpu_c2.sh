python -m pytest tests/test_gpu_nms.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for sm in 1 0; do for reg in typical dense; do
DBL_DF_SMEM=$sm python bench.py --workload decode_nms --regime $reg --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('smem=$sm $reg', d['value'], d['ms_per_step'], 'filter us', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac'], 'nms ms', d['nms']['ms'])"
done; done
