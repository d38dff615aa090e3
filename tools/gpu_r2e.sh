#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
bash tools/gpu_variants.sh r2e
timeout 2400 python tools/run_reference_conformance.py --scale 1.0 --jobs 14 --log gpurun_out/r2e_reference_conformance_full.log > gpurun_out/r2e_refconf.out 2>&1; echo "reference conformance rc=$?"; tail -2 gpurun_out/r2e_refconf.out
grep "^FAIL" gpurun_out/r2e_refconf.out | head
