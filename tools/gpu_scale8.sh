#!/bin/bash
# 8-GPU scaling check of the bench line (run with gpurun --gpus 8): one rank per GPU under torchrun, as the driver launches it.
n=${1:-8}; shift
mkdir -p gpurun_out
nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))"; nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n --no-configs --no-cpu-baseline --steps 10 --streams 4,16 "$@" > gpurun_out/r2_scale${n}.json 2> gpurun_out/r2_scale${n}.err; echo "rc=$?"; tail -3 gpurun_out/r2_scale${n}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_scale${n}.json') if l.startswith('{')][-1])
    print('N=${n} value', d['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e', d['e2e']['value'], d['e2e']['streams'], 'e2e_rdram', d['e2e_rdram']['value'])
    print('host', d['e2e']['command_thread_us_per_frame_by_streams'], 'cores/rank', d['config']['host_cores_per_rank'])
except Exception as e: print('no line', e)
PY
