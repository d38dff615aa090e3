#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/sweep.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/sweep.txt
run() { timeout 280 python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-parity "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'value', d['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e', d['e2e']['value'], d['e2e']['streams'], d['e2e']['command_thread_us_per_frame'])" >> gpurun_out/sweep.txt; }
run --streams 8,12,16
run --streams 8,16 --e2e-depth 5
run --streams 8 --e2e-depth 2
cat gpurun_out/sweep.txt
