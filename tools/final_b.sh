#!/bin/bash
# Final evidence batch of round 2 (one GPU, under gpurun): default bench line, the driver's command shape (K = 20, W = 5),
# launch list, GPU suite.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 python bench.py > gpurun_out/r02_bench_large_A_n1.json 2> gpurun_out/r02_bench_large_A_n1.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-parity > gpurun_out/r02_bench_large_A_n1_k20.json 2>/dev/null
$NCU --metrics gpu__time_duration.sum -k regex:"k_|mcgra" -c 400 --csv --log-file gpurun_out/r02_large_A_launches.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/r02_prof_a.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -s > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
python - <<'P'
import json
for f in ("r02_bench_large_A_n1.json", "r02_bench_large_A_n1_k20.json"):
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["steps"])
P
