#!/bin/bash
# tests (optional: SKIP_TESTS=1) + bench + launch list + full ncu capture of the tile kernels.  $1 = tag.
tag=${1:-q}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
fi
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --streams ${STREAMS:-1,8} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'streams', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e streams', d['e2e']['streams'])
print('kernel_us/frame', {k: round(v['total_ms']/d['steps']*1000,1) for k,v in d['kernel_ms'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 20 --warmup 3 --small --no-parity --no-cpu-baseline --streams 1 > gpurun_out/${tag}_ncu2.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-k_tiles}" --launch-skip ${SKIP:-1} -c ${COUNT:-1} -f -o gpurun_out/${tag}_full \
  python bench.py --steps 2 --warmup 1 --small --no-parity --no-cpu-baseline --streams 1 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu full rc=$?"
