#!/bin/bash
# (1) staging-layout fix (product) vs phase D by survivor index (variant dsurv); (2) the C4-ii kernels under ncu.
tag=${1:-r2o}
mkdir -p gpurun_out
bash tools/gpu_variants.sh $tag --streams 1,4 --steps 8
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'k_upscale|k_vi_downscale|k_span_bin<\(bool\)1>|k_tiles.*\(bool\)0, \(bool\)1>' --launch-skip 12 -c 10 -f -o gpurun_out/${tag}_c4ii \
  python bench.py --steps 1 --warmup 1 --small --no-parity --no-cpu-baseline --streams 1 --configs c4ii > gpurun_out/${tag}_ncu_c4ii.log 2>&1; echo "ncu c4ii rc=$?"; tail -2 gpurun_out/${tag}_ncu_c4ii.log | cut -c1-300
