cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final.json"))
print("N=1", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["roofline"]["kernels_ms_per_step"], d["roofline"].get("alp_path"), {k:(round(v["value"]), round(v["ms_per_step"],3)) for k,v in d["north_star_runs"].items()}, d["cpu_baseline"]["value"], d["clocks"], d["gpu_launches"])
PY
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-north-star-runs"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv $B > gpurun_out/r2_launches_final.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
