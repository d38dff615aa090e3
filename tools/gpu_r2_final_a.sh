#!/bin/bash
# Round-end set, part A: full GPU suite, smoke, default bench (all configs + cpu_baseline), reference arm, launch list,
# one frame's kernels under ncu --set full with source.  (Part B: tools/gpu_r2_final_b.sh.)
tag=${1:-r2z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
( time timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ); echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --small --no-parity --no-cpu-baseline --no-configs --streams 1 > gpurun_out/${tag}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 24 -c 8 -f -o gpurun_out/${tag}_all \
  python bench.py --steps 1 --warmup 1 --small --no-parity --no-cpu-baseline --no-configs --streams 1 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_rdram', d['e2e_rdram']['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, d['e2e']['streams'], d['roofline']['frac'], d.get('cpu_baseline'), d['clocks'])
for k,v in d.get('configs',{}).items(): print(k, v.get('frames_per_s'), v.get('e2e',{}).get('value'), v.get('roofline'))
r=json.loads(open('gpurun_out/${tag}_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['cpu_baseline'])
PY
