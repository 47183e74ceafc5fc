set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/r2n_bench1.json 2> gpurun_out/r2n_bench1.err
tail -c 600 gpurun_out/r2n_bench1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2n_bench1.json"))
print("N=1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernels_ms_per_step"], {k:v["value"] for k,v in d["north_star_runs"].items()}, d["cpu_baseline"]["value"], d["clocks"])
PY
