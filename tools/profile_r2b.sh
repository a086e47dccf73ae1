#!/bin/bash
# Round-2 final evidence batch (one GPU, under gpurun).  Outputs -> gpurun_out/, summaries copied to profiles/.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# default bench line (n = 65 536, Profile A) and the other shapes
timeout 900 python bench.py > gpurun_out/r02_bench_large_A_n1.json 2> gpurun_out/r02_bench_large_A_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 600 python bench.py --workload pubmed --profile B --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_pubmed_B.json 2> gpurun_out/r02_bench_pubmed_B.err
timeout 600 python bench.py --density 1 --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02_bench_large_A_density1.json 2> gpurun_out/r02_bench_density1.err
timeout 600 python bench.py --workload cora --steps 200 --warmup 20 --no-cpu --no-parity > gpurun_out/r02_bench_cora_A.json 2> gpurun_out/r02_bench_cora.err
# launch list (serialised, cold-cache: shares of the step, not absolute times)
$NCU --metrics gpu__time_duration.sum -k regex:"k_|mcgra" -c 400 --csv --log-file gpurun_out/r02_large_A_launches.csv \
  python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/r02_prof_a.log 2>&1
# full captures of the two new streaming kernels (steady-state launches)
$NCU --set full --import-source on -k regex:k_fold_rs -s 2 -c 1 -o gpurun_out/r02_fold_rs -f \
  python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_elem_rs -s 2 -c 1 -o gpurun_out/r02_elem_rs -f \
  python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
# compute-sanitizer over parity cases that run the new kernels (multi-tile, capped grids, KDE, MSE goldens)
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_attack.py -q -m gpu \
  -k "fold_persistent or elem_persistent or (test_attack_matches_reference_golden and (mse_A_n150 or kde_n90 or hsic_B_n150))" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity_sizes.py -q -m gpu \
  -k "test_multi_tile_measures_match_oracle and (HSIC or KDE)" >> gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck multi-tile rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -3 gpurun_out/r02_sanitizer_memcheck.log
