#!/bin/bash
# Round-end set, part C (after the last kernel change of the round): full GPU suite again, then racecheck + memcheck over
# the deferred tile kernel's new shared-memory protocol (survivor listing, bank-per-lane mask staging).
tag=${1:-r2z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${tag}c_pytest.log
log=gpurun_out/${tag}c_sanitizer.log; : > $log
for tool in racecheck memcheck; do
  echo "=== $tool: C2 + C3 frames (deferred exact kernel, fill pass, in-order sprite kernel, VI)" >> $log
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/c2_few_frames.py 1200 2 >> $log 2>&1; echo "rc=$?" >> $log
  for s in depth-compare-opaque interpolation-color-depth rasterization-many-primitives; do
    echo "=== $tool: forced deferred generic kernel, reference suite $s" >> $log
    RDP_B200_FORCE_DEFERRED=1 timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 oracle/_ref/rdp-conformance-b200 --suite $s --range 0 3 2>&1 | grep -v "Skipping suite" | tail -6 >> $log; echo "rc=${PIPESTATUS[0]}" >> $log
  done
done
grep "rc=" $log | sort | uniq -c; grep -i "ERROR SUMMARY\|RACECHECK SUMMARY" $log | sort | uniq -c | head
