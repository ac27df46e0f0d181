#!/bin/bash
set -u
O=gpurun_out/r02n2
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-gather --no-strong > $O/bench_n2_nogather.json 2> $O/bench_n2_nogather.err; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/bench_n2_ref.json 2> $O/bench_n2_ref.err; echo "rc=$?"
tail -3 $O/bench_n2.err
python - <<PY
import json
for f in ('bench_n2.json','bench_n2_nogather.json'):
    d=json.loads(open('$O/'+f).read().strip().splitlines()[-1])
    print(f, 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('repetitions_ms_per_step'), 'gather', d['e2e']['gather'], 'numa', d['e2e'].get('numa'))
    print('   strong', d.get('strong'))
PY
head -c 600 $O/bench_n2_ref.json
