#!/bin/bash
tag=${1:-b}; shift
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --streams ${STREAMS:-1,8} "$@" > gpurun_out/${tag}_bench.json 2>gpurun_out/${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], {k:v["frames_per_s"] for k,v in d["streams"].items()}, "e2e", d["e2e"]["value"], d["e2e"]["streams"], d["e2e"]["command_thread_us_per_frame"])
PY
