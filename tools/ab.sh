#!/bin/bash
# A/B helper for gpurun: tools/ab.sh <tag> [ENV=val ...] -- runs a short det bench (no CPU baseline) and prints tag, value, e2e
tag=$1; shift
env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${AB_ARGS} 2> gpurun_out/ab_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('$tag', 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'ms=%.3f'%d['ms_per_step'], 'top=%s %.0fus frac=%s'%(r['kernel'], r['avg_launch_us'], r['frac']))
" || tail -5 gpurun_out/ab_$tag.err
