#!/bin/bash
# Round 2, first GPU call: Vulkan ICD probe, full GPU suite, smoke, default bench.
tag=${1:-r2a}
mkdir -p gpurun_out
{
  echo "== ls /usr/share/vulkan/icd.d /etc/vulkan/icd.d"; ls -la /usr/share/vulkan/icd.d /etc/vulkan/icd.d 2>&1
  echo "== find nvidia_icd / libvulkan / libGLX_nvidia"; find / -xdev \( -name 'nvidia_icd*.json' -o -name 'libvulkan.so*' -o -name 'libGLX_nvidia.so*' -o -name 'libnvidia-vulkan*' \) 2>/dev/null | head -20
  echo "== vulkaninfo"; which vulkaninfo && vulkaninfo --summary 2>&1 | head -30
  echo "== nvidia-smi"; nvidia-smi -L; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv
  echo "== cpus"; nproc; lscpu | head -20; python -c "import os; print(sorted(os.sched_getaffinity(0)))"
} > gpurun_out/${tag}_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_rdram', d['e2e_rdram']['value'], {k:v['frames_per_s'] for k,v in d['streams'].items()}, d['e2e']['streams'])
    print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_us'], 'kernels', {k:v['us_per_frame'] for k,v in d['kernel_ms'].items()})
    for k,v in d.get('configs',{}).items(): print(k, v.get('frames_per_s'), v.get('e2e',{}).get('value'), v.get('kernel_us_per_frame'), v.get('parity'))
    print(d.get('cpu_baseline'), d.get('clocks'))
except Exception as e: print('no bench line', e)
PY
