#!/bin/bash
# Quick relative bench of the kernel variants under parallel-rdp_b200/variants/ against the product library.
# usage: gpu_variants.sh <tag> [bench args]      (defaults: --streams 1,4,16 --steps 8)
tag=${1:-var}; shift
mkdir -p gpurun_out
args="${@:---streams 1,4,16 --steps 8}"
for lib in parallel-rdp_b200/librdp_b200.so parallel-rdp_b200/variants/librdp_b200_*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  RDP_B200_LIBRARY=$PWD/$lib timeout 300 python bench.py --no-configs --no-cpu-baseline --small --parity-frames 2 $args > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_${name}.json').read().strip().splitlines()[-1])
    print('${name}', 'value', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e', d['e2e']['streams'], 'tiles us', d['roofline']['avg_launch_us'], d['config']['parity'][:20])
except Exception as e: print('${name}', 'FAILED', e, open('gpurun_out/${tag}_${name}.err').read()[-300:])
PY
done
