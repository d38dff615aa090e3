#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_reference_harness.py -m gpu -x -q ) > gpurun_out/r2f_harness_pytest.log 2>&1; tail -4 gpurun_out/r2f_harness_pytest.log
timeout 1200 python tools/run_reference_conformance.py --scale 1.0 --jobs 3 --skip-passed-in profiles/r2e_reference_conformance_full_depth_part1.log --log gpurun_out/r2f_reference_conformance_part2.log > gpurun_out/r2f_refconf.out 2>&1; echo "part2 rc=$?"; tail -2 gpurun_out/r2f_refconf.out; grep "^FAIL" gpurun_out/r2f_refconf.out | head
for lock in 1 0; do
  RDP_B200_SUBMIT_LOCK=$lock timeout 300 python bench.py --no-configs --no-cpu-baseline --small --no-parity --streams 4,16 --steps 8 > gpurun_out/r2f_lock$lock.json 2>gpurun_out/r2f_lock$lock.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2f_lock$lock.json').read().strip().splitlines()[-1]); print('submit lock $lock: value', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'e2e', d['e2e']['streams'], d['e2e']['command_thread_us_per_frame_by_streams'], 'e2e_rdram', d['e2e_rdram']['value'])"
done
