#!/bin/bash
# ensemble with all terms staged at once: tests, e2e phases, and an ncu --set full capture of the two end-to-end kernels
set -x
timeout 200 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_attack.py -q -m gpu -x 2>&1 | tail -3
timeout 200 python tools/e2e_breakdown.py large 20 > gpurun_out/r02_e2e_breakdown_large20.log 2>&1; tail -3 gpurun_out/r02_e2e_breakdown_large20.log
timeout 200 ncu --clock-control none --set full --import-source on -k regex:"k_rank_negatives_tab|k_ensemble_sym" -c 2 -o gpurun_out/r02_e2e_kernels -f python tools/e2e_breakdown.py large 20 > /dev/null 2>&1
ls -la gpurun_out/r02_e2e_kernels.ncu-rep
