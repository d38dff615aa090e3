#!/bin/bash
# The default bench line exactly as the driver launches it on N GPUs (all configs + C5 farm), to check the multi-rank paths.
n=${1:-2}
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $n --steps 20 --warmup 3 ) > gpurun_out/r2_default_n${n}.json 2> gpurun_out/r2_default_n${n}.err; echo "rc=$?"; tail -6 gpurun_out/r2_default_n${n}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_default_n${n}.json') if l.startswith('{')][-1])
    print('N=${n} value', d['value'], 'e2e', d['e2e']['value'], 'e2e_rdram', d['e2e_rdram']['value'], d['clocks'])
    for k,v in d.get('configs',{}).items(): print(k, v.get('frames_per_s'), v.get('e2e',{}).get('value'))
except Exception as e: print('no line', e)
PY
