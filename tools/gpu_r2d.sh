#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2d_smoke.log
bash tools/gpu_variants.sh r2d
for t in 8 16; do
  timeout 300 python bench.py --no-configs --no-cpu-baseline --small --no-parity --streams 4,8,16 --steps 8 --e2e-threads $t > gpurun_out/r2d_thr$t.json 2>gpurun_out/r2d_thr$t.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2d_thr$t.json').read().strip().splitlines()[-1]); print('threads $t e2e', d['e2e']['streams'], d['e2e']['command_thread_us_per_frame_by_streams'])"
done
timeout 300 python bench.py --no-configs --no-cpu-baseline --small --no-parity --streams 4,8,16 --steps 8 --e2e-threads 8 --e2e-depth 5 > gpurun_out/r2d_depth5.json 2>gpurun_out/r2d_depth5.err
python -c "
import json; d=json.loads(open('gpurun_out/r2d_depth5.json').read().strip().splitlines()[-1]); print('threads 8 depth 5 e2e', d['e2e']['streams'])"
PARALLEL_RDP_BENCH=1 timeout 300 python bench.py --no-configs --no-cpu-baseline --small --no-parity --streams 16 --steps 8 > gpurun_out/r2d_hostcalls.json 2> gpurun_out/r2d_hostcalls.err
grep "host:" gpurun_out/r2d_hostcalls.err | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tiles --launch-skip 6 -c 1 -f -o gpurun_out/r2d_tiles \
  python bench.py --steps 1 --warmup 1 --small --no-parity --no-cpu-baseline --no-configs --streams 1 > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2d_tiles.ncu-rep
