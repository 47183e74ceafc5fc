cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or golden or graphed or end_to_end or full_size" 2>&1 | tail -3
for tool in racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke OK' gpurun_out/r2_sanitizer_$tool.log | tr '\n' ' ')"
done
timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-north-star-runs > gpurun_out/r2y.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2y.json')); print('N=1', round(d['value']), round(d['ms_per_step'],4), d['roofline']['kernels_ms_per_step'])"
