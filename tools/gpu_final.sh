#!/bin/bash
# Round-end set: full GPU suite, smoke, default bench (with cpu_baseline), reference arm, launch list, full ncu captures.
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 20 --warmup 3 --small --no-parity --no-cpu-baseline --streams 1 > gpurun_out/${tag}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 16 -c 8 -f -o gpurun_out/${tag}_all \
  python bench.py --steps 2 --warmup 1 --small --no-parity --no-cpu-baseline --streams 1 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, d['e2e']['streams'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'])
r=json.loads(open('gpurun_out/${tag}_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['cpu_baseline'])
PY
