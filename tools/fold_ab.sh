# same-box A/B of the fold engines at the bench shape (n = 65 536, Profile A): MCGRA_ENGINES="1:<engine>"
for e in "1:3" "1:2" "1:3" "1:2"; do
MCGRA_ENGINES="$e" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/fold_ab.json 2>gpurun_out/fold_ab.err
python - "$e" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/fold_ab.json"))
    pk={k["kernel"]:k["ms"] for k in d["roofline"]["per_kernel"]}
    print("engines", sys.argv[1], ": it/s", round(d["value"],2), "ms", round(d["ms_per_step"],2), "fold", round(pk["mcgra_fold_adam"],3), "loss", d["loss_first_last"][1])
except Exception as ex:
    print("engines", sys.argv[1], "FAILED", ex, open("gpurun_out/fold_ab.err").read()[-300:])
PY
done
