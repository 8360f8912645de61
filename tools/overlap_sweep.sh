#!/bin/bash
# ORBX_OPT_OVERLAP on/off and FAST resident-warp caps, kernel-only bench (run on the GPU box)
TAG=${1:-ovl}
{
python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_stream.py -x -q 2>&1 | tail -2
for args in "--no-overlap" "" "--fast-ctas 8" "--fast-ctas 12" "--fast-ctas 16" "--fast-ctas 20"; do
  for rep in 1 2; do
  python bench.py --steps 20 --warmup 3 --kernels-only $args 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$args]', '%.0f fps  %.4f ms/step' % (d['value'], d['ms_per_step']), ' '.join('%s %.3f' % (k.replace('k_', ''), v['ms_per_step']) for k, v in d['kernels'].items()))"
  done
done
} 2>&1 | tee gpurun_out/${TAG}_overlap.log
