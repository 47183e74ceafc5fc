set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
for v in "3 2" "4 2" "3 6" "4 6" "3 3"; do set -- $v
  PSAM_TC_STAGES=$1 PSAM_BW_CTAS=$2 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_s$1_c$2.json 2> gpurun_out/r2_bench_s$1_c$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_s$1_c$2.json"))
print("stages $1 ctas $2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernels_ms_per_step"])
PY
done
