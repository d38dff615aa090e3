#!/bin/bash
# BASELINE C4-ii: C2 with 4x internal upscaling, scan-out with 0 and 2 downscale steps (and 2x for reference).
mkdir -p gpurun_out
for cfg in "4 0" "4 2" "2 1"; do
  set -- $cfg
  timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --small --streams 1,4 --upscale $1 --downscale-steps $2 > gpurun_out/c4_up$1_ds$2.json 2> gpurun_out/c4_up$1_ds$2.err
  echo "upscale $1 downscale $2 rc=$?"; tail -3 gpurun_out/c4_up$1_ds$2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c4_up$1_ds$2.json').read().strip().splitlines()[-1])
    print('  value', d['value'], 'e2e', d['e2e']['value'], 'streams', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'kernel_us/frame', {k: round(v['total_ms']/d['steps']*1000,1) for k,v in d['kernel_ms'].items()}, d['config']['parity'])
except Exception as e: print('  no result', e)
PY
done
