#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the hot kernels: C2 / C3 frames through the API and a spread of the reference's
# conformance suites (in-order, deferred, fill, copy, texture loads, VI), each a few iterations.  Log -> gpurun_out/<tag>_sanitizer.log
tag=${1:-r2}
log=gpurun_out/${tag}_sanitizer.log
mkdir -p gpurun_out; : > $log
SUITES="fill-16 tex-rect interpolation-color-texture-rgba16 depth-compare-opaque copy-16bpp-fb16 texture-load-tile-16 texture-load-block-16 texture-load-tlut-8 color-on-coverage rasterization-many-primitives"
for tool in memcheck racecheck; do
  echo "=== $tool: C2 + C3 frames (deferred exact kernel, fill pass, in-order sprite kernel, VI)" >> $log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/c2_few_frames.py 1200 2 >> $log 2>&1; echo "rc=$?" >> $log
  for s in $SUITES; do
    echo "=== $tool: reference suite $s" >> $log
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 oracle/_ref/rdp-conformance-b200 --suite $s --range 0 3 2>&1 | grep -v "Skipping suite" | tail -6 >> $log; echo "rc=${PIPESTATUS[0]}" >> $log
  done
  echo "=== $tool: forced deferred generic kernel, depth-compare-opaque + interpolation-color-depth" >> $log
  for s in depth-compare-opaque interpolation-color-depth; do
    RDP_B200_FORCE_DEFERRED=1 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 oracle/_ref/rdp-conformance-b200 --suite $s --range 0 3 2>&1 | grep -v "Skipping suite" | tail -6 >> $log; echo "rc=${PIPESTATUS[0]}" >> $log
  done
  echo "=== $tool: VI suites" >> $log
  for s in $(grep "^vi " tests/golden/reference_ctest_list.txt | head -3 | awk '{print $2}'); do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 oracle/_ref/vi-conformance-b200 --suite $s --range 0 2 2>&1 | grep -v "Skipping suite" | tail -6 >> $log; echo "rc=${PIPESTATUS[0]}" >> $log
  done
done
grep -c "rc=0" $log; grep "rc=" $log | sort | uniq -c; grep -i "ERROR SUMMARY\|RACECHECK SUMMARY" $log | sort | uniq -c | head
