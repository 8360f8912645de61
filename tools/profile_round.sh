#!/bin/bash
# Runs on the GPU box (under gpurun): un-profiled bench, ncu launch list of one step, ncu --set full of the step's kernels.
# The kernels outside the step chain: tools/profile_extra.sh (its own gpurun call: gpurun_out/ is limited to 64 MiB per call).
# usage: tools/profile_round.sh <tag>
set -u
TAG=${1:-r02}
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || exit 1
CMD="python bench.py --steps 1 --warmup 3 --kernels-only --batch 32"
$CMD > $O/${TAG}_plain.log 2>&1 || exit 1
# 15 launches per step (7 resize, fast, quadtree, describe with the Gaussian fused, filter, 2x2 match): any 15 consecutive launches past the warm-up hold one of each
ncu --metrics gpu__time_duration.sum --clock-control none -s 46 -c 15 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -s 46 -c 15 -o $O/${TAG}_full -f $CMD > $O/${TAG}_ncu_full.log 2>&1
tail -2 $O/${TAG}_ncu_full.log
