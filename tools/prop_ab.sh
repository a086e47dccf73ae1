# ablation of the fp16x2 propagate engine at the bench shape: MCGRA_ENGINES="0:<100 + bits>" (1 = no flush, 2 = no MMAs)
for e in "0:100" "0:101" "0:102" "0:103"; do
MCGRA_ENGINES="$e" timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/prop_ab.json 2>gpurun_out/prop_ab.err
python - "$e" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/prop_ab.json"))
    pk={k["kernel"]:k["ms"] for k in d["roofline"]["per_kernel"]}
    print("engines", sys.argv[1], ": prop32", round(pk["propagate32"],3), "prop16", round(pk["propagate16"],3), "fold", round(pk["mcgra_fold_adam"],3))
except Exception as ex:
    print("engines", sys.argv[1], "FAILED", ex, open("gpurun_out/prop_ab.err").read()[-300:])
PY
done
