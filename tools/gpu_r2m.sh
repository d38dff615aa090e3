#!/bin/bash
# List-kernel A/B: product with and without RDP_B200_NO_TILE_LISTS, then the deferred parity subset.
tag=${1:-r2m}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-cpu-baseline --small --parity-frames 2 --streams 1,4 --steps 8 > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e', d['e2e']['streams'], 'tiles us', d['roofline']['avg_launch_us'], {k:v['us_per_frame'] for k,v in d['kernel_ms'].items()}, d['config']['parity'][:20])
except Exception as e: print('$name', 'FAILED', e, open('gpurun_out/${tag}_$name.err').read()[-400:])
PY
}
run lists X=1
run nolists RDP_B200_NO_TILE_LISTS=1
timeout 600 python -m pytest tests/test_gpu_conformance.py tests/test_workloads.py -m gpu -x -q -k "deferred or many_overlapping or workload" > gpurun_out/${tag}_pytest_deferred.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_deferred.log
