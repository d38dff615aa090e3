#!/bin/bash
# Quick 1-GPU check of the e2e legs (host CPU per frame, cores busy).  usage: gpu_quick_e2e.sh <tag> [bench args]
tag=${1:-q}; shift
mkdir -p gpurun_out
timeout 300 python bench.py --no-configs --no-cpu-baseline --small --parity-frames 2 --streams 1,4,16 --steps 8 "$@" > gpurun_out/${tag}_quick.json 2>gpurun_out/${tag}_quick.err; echo "rc=$?"; tail -2 gpurun_out/${tag}_quick.err
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_quick.json').read().strip().splitlines()[-1]); h=d['e2e']['command_thread_us_per_frame_by_streams']; print('value', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e', d['e2e']['streams'], 'cores busy', {s:h[s]['cores_busy'] for s in h}, 'cpu us/frame', {s:h[s]['process_cpu_us'] for s in h}, 'e2e_rdram', d['e2e_rdram']['value'], d['config']['parity'][:30]); print(h)"
