#!/bin/bash
# Runs ON THE GPU BOX: one bench line + the per-launch table, printed by family and for the rows matching $1 (regex on the launch name)
O=gpurun_out/kt
mkdir -p $O
timeout 300 python bench.py --no-cpu-baseline --steps 20 --kernel-table $O/kt.json > $O/bench.json 2> $O/bench.err || { tail -5 $O/bench.err; exit 1; }
python - "$1" <<'PY'
import json,re,sys
d=json.loads(open('gpurun_out/kt/bench.json').read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
t=json.load(open('gpurun_out/kt/kt.json'))
print('serialised', round(sum(f['ms'] for f in t['families']),3))
for f in t['families'][:10]: print(' ', f['kern'], f['launches'], round(f['ms'],3))
pat=sys.argv[1] if len(sys.argv)>1 and sys.argv[1] else None
if pat:
    for r in t['launches']:
        if re.search(pat, r['what']) or re.search(pat, r['kern']): print(r['what'], r['kern'], round(r['ms']*1e3,1))
PY
