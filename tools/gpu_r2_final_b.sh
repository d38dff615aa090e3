#!/bin/bash
# Round-end set, part B: parity beyond the reference's ctest depth (forced-deferred generic kernel through the reference's
# suites, 100 benchmark frames), and the C4-ii kernels (the one byte-heavy config) under ncu.
tag=${1:-r2z}
mkdir -p gpurun_out
RDP_B200_FORCE_DEFERRED=1 timeout 900 python tools/run_reference_conformance.py --single-process 60 --log gpurun_out/${tag}_reference_conformance_forced_deferred.log > gpurun_out/${tag}_refconf_fd.out 2>&1; echo "forced-deferred conformance rc=$?"; tail -1 gpurun_out/${tag}_refconf_fd.out
timeout 600 python tools/c2_few_frames.py 5000 100 30 > gpurun_out/${tag}_c2_c3_100_frames.log 2>&1; echo "100 frames rc=$?"; cat gpurun_out/${tag}_c2_c3_100_frames.log
timeout 600 ncu --set full --clock-control none -k regex:"k_tiles|k_upscale|k_span_bin|k_vi" --launch-skip 40 -c 14 -f -o gpurun_out/${tag}_c4ii \
  python bench.py --steps 1 --warmup 1 --small --no-parity --no-cpu-baseline --streams 1 --configs c4ii > gpurun_out/${tag}_ncu_c4ii.log 2>&1; echo "ncu c4ii rc=$?"
