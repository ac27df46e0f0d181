#!/bin/bash
# e2e variability probe: heat the GPU for a minute (what a test suite does before the bench), then many e2e repetitions with the clocks / power /
# throttle reasons sampled by ONE persistent nvidia-smi process
python bench.py --quick --steps ${1:-15000} > /dev/null 2>&1
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.hw_slowdown --format=csv,noheader -lms 50 > /tmp/smi.log 2>&1 &
SMI=$!
DBL_NO_CLK=1 DBL_E2E_DIAG=1 DBL_E2E_REPS=25 python bench.py --no-cpu-baseline 2>&1 >/tmp/line.json | grep "e2e rep"
kill $SMI
python - <<'PY'
import collections
rows=[l.strip().split(', ') for l in open('/tmp/smi.log') if l.count(',')>=6]
clk=[int(r[0].split()[0]) for r in rows]; pw=[float(r[1].split()[0]) for r in rows]; t=[int(r[2]) for r in rows]
print('samples', len(rows), 'sm MHz min/median/max', min(clk), sorted(clk)[len(clk)//2], max(clk), 'power max', max(pw), 'temp max', max(t))
print('reasons active:', {n: sum(1 for r in rows if r[3+i].lower().startswith('active')) for i,n in enumerate(['sw_power_cap','sw_thermal','hw_thermal','hw_slowdown'])})
low=[(c,p_,tt) for c,p_,tt in zip(clk,pw,t) if c<1900]
print('samples below 1900 MHz:', len(low), low[:10])
PY
