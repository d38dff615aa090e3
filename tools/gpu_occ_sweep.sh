#!/bin/bash
mkdir -p gpurun_out
for n in 8 6 5 4 3; do
  RDP_B200_TILE_BLOCKS_PER_SM=$n timeout 300 python bench.py --steps 150 --warmup 3 --no-cpu-baseline --no-parity --streams 1,8,16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('blocks/SM $n', 'value', d['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e', d['e2e']['streams'])"
done
