set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r2f_bench1.json 2> gpurun_out/r2f_bench1.err
tail -c 3000 gpurun_out/r2f_bench1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2f_bench1.json"))
print("N=1", d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["kernels_ms_per_step"], d["north_star_runs"], d["cpu_baseline"])
PY
