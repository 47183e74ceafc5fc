cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_driver_ref.json 2> gpurun_out/r2_driver_ref.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_driver_ours.json 2> gpurun_out/r2_driver_ours.err
python - <<PY
import json
r=json.load(open("gpurun_out/r2_driver_ref.json")); d=json.load(open("gpurun_out/r2_driver_ours.json"))
print("ref", r["value"], r["cpu_baseline"], r.get("ms_per_step"))
print("ours", round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
print("ratio e2e", d["e2e"]["value"]/r["value"], "value", d["value"]/r["value"])
PY
