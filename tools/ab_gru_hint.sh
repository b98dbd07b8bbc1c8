for h in 0 1 0 1; do
DEVO_GRU_L2_HINT=$h timeout 300 python bench.py 2>/dev/null > gpurun_out/abg_$h.json
python - <<PY
import json
d=json.load(open("gpurun_out/abg_$h.json"))
print("gru hint=$h value",d["value"],"warm",d["config"]["value_l2_warm"],"e2e",d["e2e"]["value"],"gru",d["per_op_us"]["update_operator"],"status",d["config"]["ba_status"])
PY
done
