#!/bin/bash
# ncu evidence for profiles/: launch list (shares) + one --set full capture of the hot kernels.  Run under gpurun.
# usage: bash profiles/ncu_capture.sh <workload> <tag> [skip-launch-list]
mkdir -p gpurun_out
W=${1:-pubmed}; TAG=${2:-v}
if [ -z "$3" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${W}_$TAG.csv \
    python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench_$W.log 2>&1
fi
ncu --set full --clock-control none --import-source on -k regex:'k_fold|k_propagate|k_pairs|k_elem' -s 14 -c 7 \
    -o gpurun_out/prof_${W}_$TAG -f python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_$W.log 2>&1
ls -la gpurun_out | tail -5
