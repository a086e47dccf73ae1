#!/bin/bash
# Round-2 profiling batch (run under gpurun, one GPU).  Outputs go to gpurun_out/; summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch lists (serialised, cold-cache: shares of the step, not absolute times)
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r02_large_A_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_prof_a.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_pubmed_B_launches.csv \
  python bench.py --workload pubmed --profile B --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_prof_b.log 2>&1
# full captures of the kernels of interest (one launch each)
$NCU --set full --import-source on -k regex:k_pairs_tc -c 1 -o gpurun_out/r02_pairs_tc -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_fold_tc -c 1 -o gpurun_out/r02_fold_tc -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_gemm3 -s 2 -c 1 -o gpurun_out/r02_gemm3 -f \
  python bench.py --workload pubmed --profile B --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
# compute-sanitizer over the small parity cases (n = 150 single/two tile rows, n = 700 multi-tile incl. the TMA GEMM)
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_attack.py -q -m gpu \
  -k "test_attack_matches_reference_golden and (mse_A_n150 or hsic_B_n150 or kl_all_n90 or mse_budget_n150)" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity_sizes.py -q -m gpu \
  -k "test_multi_tile_measures_match_oracle and HSIC" >> gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck multi-tile rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_attack.py -q -m gpu \
  -k "test_attack_matches_reference_golden and (mse_A_n150 or hsic_B_n150)" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -5 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
