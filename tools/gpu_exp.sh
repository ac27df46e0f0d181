for v in 32 16; do
DBL_FUSE_BN_MAXC=$v python bench.py --no-cpu-baseline --steps 10 --kernel-table gpurun_out/kt_bn$v.json 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('FUSE_BN_MAXC=$v value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['repetitions_ms_per_step'])"
python - <<PY
import json
d=json.load(open('gpurun_out/kt_bn$v.json'))
for r in d['launches']:
    if r['what'].startswith(('model.2','model.4','model.3','model.5')): print('   ', r['what'], r['kern'], round(r['ms']*1e3,1))
PY
done
