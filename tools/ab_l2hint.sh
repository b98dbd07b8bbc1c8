for h in 0 1 2 0 1 2; do
DEVO_CORR_L2_HINT=$h timeout 300 python bench.py 2>/dev/null > gpurun_out/ab_$h.json
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$h.json"))
print("hint=$h value",d["value"],"warm",d["config"]["value_l2_warm"],"e2e",d["e2e"]["value"],"corr",d["per_op_us"]["corr_lookup"],"gru",d["per_op_us"]["update_operator"],"in_step",d["in_step_us"]["corr_lookup"],d["in_step_us"]["update_operator"])
PY
done
