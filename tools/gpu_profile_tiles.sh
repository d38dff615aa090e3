#!/bin/bash
# ncu capture of the main (C2 triangles) tile kernel + a launch list.  $1 = tag for the output names.
tag=${1:-prof}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tiles --launch-skip 2 -c 2 -f -o gpurun_out/${tag}_tiles \
  python bench.py --steps 2 --warmup 1 --small --no-parity --no-cpu-baseline --streams 1 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 20 --warmup 3 --small --no-parity --no-cpu-baseline --streams 1 > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu launches rc=$?"
