#!/bin/bash
# quick GPU pass: parity tests (bounded), then the variant sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/quick_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/quick_pytest.log
tail -15 gpurun_out/quick_pytest.log
if [ -n "$1" ]; then timeout 900 bash scripts/sweep_variants.sh "$@" | tee gpurun_out/quick_sweep.log; fi
