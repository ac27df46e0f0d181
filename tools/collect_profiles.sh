#!/bin/bash
# Runs ON THE GPU BOX (gpurun): bench lines first (no profiler), then the ncu launch list and one --set full capture per kernel family.
# The .ncu-rep files are summarised on the box (tools/ncu_summary.py, tools/ncu_traffic.py) and deleted: gpurun_out/ returns at most 64 MiB.
#   bash tools/collect_profiles.sh r02
set -u
R=${1:-r02}
O=gpurun_out/$R
mkdir -p $O
python bench.py > $O/${R}_bench_n1.json 2> $O/bench_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/bench_ref.err
python bench.py --no-cpu-baseline --steps 5 --kernel-table $O/${R}_kernel_table.json > /dev/null 2>&1
python bench.py --workload decode_nms --regime typical > $O/${R}_bench_config2_typical.json 2> $O/c2t.err
python bench.py --workload decode_nms --regime dense --steps 10 > $O/${R}_bench_config2_dense.json 2> $O/c2d.err
python bench.py --scale l --batch 32 --steps 10 --no-cpu-baseline --kernel-table $O/${R}_kernel_table_l.json > $O/${R}_bench_config4_l_b32.json 2> $O/c4.err
python bench.py --imgsz 1280 --batch 16 --conf 0.001 --steps 10 --no-cpu-baseline > $O/${R}_bench_config5_1280_b16.json 2> $O/c5.err
B="python bench.py --steps 2 --warmup 3 --quick --inflight 1"
$B > $O/plain_quick.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${R}_launches.csv $B > $O/ncu_launches.log 2>&1
F="--set full --clock-control none"
T=/tmp/ncu_$R
mkdir -p $T
ncu $F -k 'regex:conv3x3_tc4|conv3x3_tc5|conv3x3s2_tc|bottleneck_tc' -c 24 -f -o $T/prof_tc4_$R $B > $O/ncu_tc4.log 2>&1
ncu $F -k regex:conv_tc_kernel -c 20 -f -o $T/prof_tc_$R $B > $O/ncu_tc.log 2>&1
ncu $F -k 'regex:dwconv_tile3|dsconv_mma|dsconv_tc' -c 20 -f -o $T/prof_dw_$R $B > $O/ncu_dw.log 2>&1
ncu $F -k 'regex:stem_pair|head_decode|sort_nms|dysample|gated_add|hg_' -c 16 -f -o $T/prof_misc_$R $B > $O/ncu_misc.log 2>&1
python tools/ncu_summary.py $T/prof_tc4_$R.ncu-rep $T/prof_tc_$R.ncu-rep $T/prof_dw_$R.ncu-rep $T/prof_misc_$R.ncu-rep > $O/${R}_ncu_table.md 2> $O/ncu_summary.err
python tools/ncu_traffic.py $T/prof_tc4_$R.ncu-rep $T/prof_tc_$R.ncu-rep $T/prof_dw_$R.ncu-rep $T/prof_misc_$R.ncu-rep > $O/${R}_ncu_traffic.json 2> $O/ncu_traffic.err
ls -la $O $T
