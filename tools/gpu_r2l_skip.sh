for n in skip8 skip10 skip12; do
RDP_B200_LIBRARY=$PWD/parallel-rdp_b200/variants/librdp_b200_$n.so timeout 300 python bench.py --no-configs --no-cpu-baseline --small --no-parity --streams 1,4 --steps 8 > gpurun_out/r2l_$n.json 2> gpurun_out/r2l_$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_$n.json').read().strip().splitlines()[-1])
print('$n', {k:v['frames_per_s'] for k,v in d['streams'].items()}, 'tiles us', d['roofline']['avg_launch_us'], d['kernel_ms'])
PY
done
