#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_reference_harness.py -m gpu -x -q ) > gpurun_out/r2g_harness_pytest.log 2>&1; tail -4 gpurun_out/r2g_harness_pytest.log | head -2
for k in 1 2 3 4 6 0; do
  RDP_B200_SUBMIT_CONCURRENCY=$k timeout 300 python bench.py --no-configs --no-cpu-baseline --small --no-parity --streams 4,16 --steps 8 > gpurun_out/r2g_gate$k.json 2>gpurun_out/r2g_gate$k.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2g_gate$k.json').read().strip().splitlines()[-1]); h=d['e2e']['command_thread_us_per_frame_by_streams']; print('gate $k: e2e', d['e2e']['streams'], 'cores busy', {s:h[s]['cores_busy'] for s in h}, 'cpu us/frame', {s:h[s]['process_cpu_us'] for s in h}, 'submit', {s:h[s]['submit_s'] for s in h}, 'e2e_rdram', d['e2e_rdram']['value'])"
done
