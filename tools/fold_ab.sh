# same-box A/B of fold variants at the bench shape (n = 65 536, Profile A): MCGRA_ENGINES="1:<engine or 1000 + experiment bits>"
for e in "1:1000" "1:1003" "1:1007" "1:1002" "1:1004" "1:1000" "1:1007"; do
MCGRA_ENGINES="$e" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/fold_ab.json 2>gpurun_out/fold_ab.err
python - "$e" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/fold_ab.json"))
    pk={k["kernel"]:k["ms"] for k in d["roofline"]["per_kernel"]}
    print("engines", sys.argv[1], ": it/s", round(d["value"],2), "ms", round(d["ms_per_step"],2), "fold", round(pk["mcgra_fold_adam"],3), "elem", round(pk["elem_stats"],3), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"))
except Exception as ex:
    print("engines", sys.argv[1], "FAILED", ex, open("gpurun_out/fold_ab.err").read()[-300:])
PY
done
