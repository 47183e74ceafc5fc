cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final.json"))
print("N=1", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["roofline"]["kernels_ms_per_step"], d["roofline"].get("alp_path"), {k:round(v["value"]) for k,v in d["north_star_runs"].items()}, d["cpu_baseline"]["value"], d["clocks"])
PY
for tool in memcheck racecheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke OK' gpurun_out/r2_sanitizer_$tool.log | tr '\n' ' ')"
done
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-north-star-runs"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv $B > gpurun_out/r2_launches_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_match_ts -s 4 -c 1 -f -o gpurun_out/r2_final_k_match_ts $B --lanes 1 --no-graphs > gpurun_out/r2_final_k_match_ts.log 2>&1
ls -la gpurun_out | tail -8
