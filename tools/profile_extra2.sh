#!/bin/bash
# ncu --set full of the kernels added late in round 2 (tools/profile_extra2.py): source import on (small report: few launches)
set -u
TAG=${1:-r02m}
O=gpurun_out
python tools/profile_extra2.py > $O/${TAG}_extra2_plain.log 2>&1 || { tail -5 $O/${TAG}_extra2_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:k_assoc_mma|k_fast_dense|k_fast_nms|k_fast_retry_list|k_fast_cells|k_fmat_hypotheses" \
    --launch-skip 0 --launch-count 24 -o $O/${TAG}_extra2 -f python tools/profile_extra2.py > $O/${TAG}_ncu_extra2.log 2>&1
tail -2 $O/${TAG}_ncu_extra2.log; ls -la $O | grep ${TAG}
