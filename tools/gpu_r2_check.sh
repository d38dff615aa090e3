#!/bin/bash
# Full GPU suite + a main-line bench (no sub-configs).  usage: gpu_r2_check.sh <tag> [extra bench args]
tag=${1:-r2x}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --no-configs --no-cpu-baseline "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_rdram', d['e2e_rdram']['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, d['e2e']['streams'])
    print('host us/frame', d['e2e']['command_thread_us_per_frame'], 'launches', d['gpu_launches'])
    print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_us'], 'kernels', {k:v['us_per_frame'] for k,v in d['kernel_ms'].items()})
except Exception as e: print('no bench line', e)
PY
