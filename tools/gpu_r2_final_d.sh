#!/bin/bash
# Round-end set, part D (static first tile adopted): full GPU suite, default bench, sanitizer over the deferred kernel,
# launch list and one frame's kernels under ncu --set full -- in order of importance (the GPU budget may cut the tail).
tag=${1:-r2y}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${tag}_pytest.log
( time timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ); echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_rdram', d['e2e_rdram']['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, d['e2e']['streams'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['clocks'])
for k,v in d.get('configs',{}).items(): print(k, v.get('frames_per_s'), v.get('e2e',{}).get('value'), v.get('roofline'))
PY
log=gpurun_out/${tag}_sanitizer.log; : > $log
for tool in racecheck memcheck; do
  echo "=== $tool: C2 + C3 frames (deferred exact kernel, fill pass, in-order sprite kernel, VI)" >> $log
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python tools/c2_few_frames.py 1200 2 >> $log 2>&1; echo "rc=$?" >> $log
  for s in depth-compare-opaque rasterization-many-primitives; do
    echo "=== $tool: forced deferred generic kernel, reference suite $s" >> $log
    RDP_B200_FORCE_DEFERRED=1 timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 oracle/_ref/rdp-conformance-b200 --suite $s --range 0 3 2>&1 | grep -v "Skipping suite" | tail -6 >> $log; echo "rc=${PIPESTATUS[0]}" >> $log
  done
done
grep "rc=" $log | sort | uniq -c; grep -i "ERROR SUMMARY\|RACECHECK SUMMARY" $log | sort | uniq -c | head
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --small --no-parity --no-cpu-baseline --no-configs --streams 1 > gpurun_out/${tag}_ncu2.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 24 -c 8 -f -o gpurun_out/${tag}_all \
  python bench.py --steps 1 --warmup 1 --small --no-parity --no-cpu-baseline --no-configs --streams 1 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
