#!/bin/bash
# One GPU-box call: deferred-kernel parity (workload equality + a spread of conformance suites), then the bench with the
# deferred tile kernel and, for comparison, without it.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_workloads.py -m gpu -x -q -k "deferred or full_size" > gpurun_out/t_workloads.log 2>&1
echo "workloads rc=$?"; tail -3 gpurun_out/t_workloads.log
timeout 900 python -m pytest tests/test_gpu_conformance.py -m gpu -q -k "deferred" > gpurun_out/t_deferred.log 2>&1
echo "deferred conformance rc=$?"; tail -5 gpurun_out/t_deferred.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_deferred.json 2> gpurun_out/bench_deferred.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_deferred.json').read().strip().splitlines()[-1])
print('deferred: value', d['value'], 'e2e', d['e2e']['value'], 'streams', d['streams'], 'kernel_ms', d['kernel_ms'], 'roof', d['roofline']['avg_launch_us'])
PY
RDP_B200_NO_DEFERRED=1 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --streams 1,8 > gpurun_out/bench_inorder.json 2> gpurun_out/bench_inorder.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_inorder.json').read().strip().splitlines()[-1])
print('in-order: value', d['value'], 'e2e', d['e2e']['value'], 'streams', d['streams'], 'kernel_ms', d['kernel_ms'], 'roof', d['roofline']['avg_launch_us'])
PY
