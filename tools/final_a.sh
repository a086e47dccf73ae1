set -x
NCU="ncu --clock-control none"
timeout 900 python bench.py > gpurun_out/r02_bench_large_A_n1.json 2> gpurun_out/r02_bench_large_A_n1.err
timeout 300 python bench.py --density 1 --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02_bench_large_A_density1.json 2>/dev/null
timeout 300 python bench.py --workload cora --steps 200 --warmup 20 --no-cpu --no-parity > gpurun_out/r02_bench_cora_A.json 2>/dev/null
$NCU --metrics gpu__time_duration.sum -k regex:"k_|mcgra" -c 400 --csv --log-file gpurun_out/r02_large_A_launches.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/r02_prof_a.log 2>&1
$NCU --set full --import-source on -k regex:k_propagate_h -s 8 -c 1 -o gpurun_out/r02_prop_h -f python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_pairs_tc -s 2 -c 1 -o gpurun_out/r02_pairs_tc -f python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
timeout 600 python -m pytest tests -q -m gpu -s > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
