#!/bin/bash
# ncu --set full of the kernels outside the per-step chain (tools/profile_extra.py): one launch of each, no source import (report size).
set -u
TAG=${1:-r02}
O=gpurun_out
python tools/profile_extra.py > $O/${TAG}_extra_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k "regex:k_match_mma|k_match_partial|k_assoc_partial|k_blur7|k_cull|k_cfast|k_cretain|k_cblur|k_fmat_score|k_resize_exact|k_describe_c|k_describe$" \
    --launch-skip 0 --launch-count 40 -o $O/${TAG}_extra -f python tools/profile_extra.py > $O/${TAG}_ncu_extra.log 2>&1
tail -2 $O/${TAG}_ncu_extra.log; ls -la $O
