# Diagnostic sweep of the e2e leg (frames in flight, copy/compute overlap, input footprint).  Run on a GPU box.
run() { timeout 280 python bench.py --steps 200 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value'], d['e2e']['value'], d['e2e']['command_thread_us_per_step'])"; }
run
run --no-parity
run --small
run --no-parity --e2e-depth 4
