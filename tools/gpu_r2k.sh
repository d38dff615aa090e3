#!/bin/bash
# Round-2 experiment batch K: product (fused survivor listing, early entry loads, reciprocal tile split) vs HEAD and
# single-change variants; L1 / shared carve-out sensitivity; deferred-kernel parity subset.
tag=${1:-r2k}
mkdir -p gpurun_out
bash tools/gpu_variants.sh $tag --streams 1,4 --steps 8
for c in 100 60; do
  RDP_B200_TILE_CARVEOUT=$c timeout 300 python bench.py --no-configs --no-cpu-baseline --small --parity-frames 2 --streams 1,4 --steps 8 > gpurun_out/${tag}_carve$c.json 2> gpurun_out/${tag}_carve$c.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_carve$c.json').read().strip().splitlines()[-1])
print('carveout $c', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'tiles us', d['roofline']['avg_launch_us'])
PY
done
RDP_B200_LIBRARY=$PWD/parallel-rdp_b200/variants/librdp_b200_cap192.so RDP_B200_TILE_CARVEOUT=58 timeout 300 python bench.py --no-configs --no-cpu-baseline --small --parity-frames 2 --streams 1,4 --steps 8 > gpurun_out/${tag}_cap192_carve58.json 2> gpurun_out/${tag}_cap192_carve58.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_cap192_carve58.json').read().strip().splitlines()[-1])
print('cap192 carveout 58', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'tiles us', d['roofline']['avg_launch_us'])
PY
timeout 600 python -m pytest tests/test_gpu_conformance.py tests/test_workloads.py -m gpu -x -q -k "deferred or many_overlapping or workload" > gpurun_out/${tag}_pytest_deferred.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_deferred.log
