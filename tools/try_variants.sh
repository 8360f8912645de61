#!/bin/bash
# Time every library build under dynamic-visual-slam_b200/lib/variants/ with the kernel-only bench (run on the GPU box; the box is scratch).
# usage: tools/try_variants.sh [bench args]
L=dynamic-visual-slam_b200/lib
for v in $L/variants/*.so; do
    cp "$v" $L/liborbx.so
    python bench.py --steps 10 --warmup 3 --kernels-only "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$(basename $v)', '%.0f fps' % d['value'], ' '.join('%s %.3f' % (k.replace('k_', ''), v['ms_per_step']) for k, v in d['kernels'].items()))"
done
