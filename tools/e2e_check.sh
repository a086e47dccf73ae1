#!/bin/bash
# e2e evidence batch: metric / ensemble tests, the phase breakdown of the public-API call, the full GPU suite.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_metrics.py -q -m gpu -x > gpurun_out/r02_metrics_tests.log 2>&1; tail -3 gpurun_out/r02_metrics_tests.log
timeout 200 python tools/e2e_breakdown.py large 20 > gpurun_out/r02_e2e_breakdown_large20.log 2>&1; tail -4 gpurun_out/r02_e2e_breakdown_large20.log
timeout 600 python -m pytest tests -q -m gpu -s > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
