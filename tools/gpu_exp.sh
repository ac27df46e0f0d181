for p in 2 3 4; do
python bench.py --no-cpu-baseline --steps 20 --inflight $p 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('inflight=$p value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
for l in 4 16; do
DBL_LANES=$l python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lanes=$l value', d['value'], 'ms', d['ms_per_step'], 'lat', d['latency_single_step_ms'])"
done
