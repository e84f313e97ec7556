#!/bin/bash
# Round profile capture (run on the GPU box through gpurun): bench lines without a profiler first, then the ncu launch
# list of the same command, then one `--set full` capture of the top kernels.  Everything lands in gpurun_out/.
set -u
R=${1:-r1}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench_rpn.json 2> $O/${R}_bench_rpn.err || exit 1
python bench.py --workload train > $O/${R}_bench_train.json 2> $O/${R}_bench_train.err || exit 1
python bench.py --workload infer > $O/${R}_bench_infer.json 2> $O/${R}_bench_infer.err || exit 1
python tools/stage_profile.py > $O/${R}_stage_profile.log 2>&1
python tools/nms_profile.py > $O/${R}_nms_phases.log 2>&1
python tools/topk_profile.py > $O/${R}_topk_phases.log 2>&1
python tools/msroialign_profile.py > $O/${R}_msroialign.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nms_keeplist|rpn_decode|topk_bucket" -s 12 -c 3 \
    -o $O/prof_${R}_proposal -f python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"roi_|sample_|region_loss" -c 8 \
    -o $O/prof_${R}_roi -f python tools/roi_one.py > $O/ncu3.log 2>&1
ls -la $O
