#!/bin/bash
# Time library builds (csrc/Makefile variant builds, lib/liborbx*.so) with the kernel-only bench and check each against the oracle
# (run on the GPU box).   usage: tools/try_variants.sh tag [variant.so ...]      default: every lib/liborbx*.so
TAG=${1:-var}; shift
L=dynamic-visual-slam_b200/lib
VARS=${@:-$L/liborbx*.so}
for v in $VARS; do
    export ORBX_LIB=$PWD/$v
    python -m pytest tests/test_gpu_parity.py -x -q -k "stages_bit_exact or awkward or retry or capacity or tie_heavy" 2>&1 | tail -1 | sed "s|^|$(basename $v) parity: |"
    for rep in 1 2; do
    python bench.py --steps 10 --warmup 3 --kernels-only 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$(basename $v)', '%.0f fps' % d['value'], ' '.join('%s %.3f' % (k.replace('k_', ''), v['ms_per_step']) for k, v in d['kernels'].items()))"
    done
done 2>&1 | tee gpurun_out/${TAG}_variants.log
