set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for v in "3 2" "3 3"; do set -- $v
  PSAM_TC_STAGES=$1 PSAM_BW_CTAS=$2 timeout 200 python tools/overlap_probe.py
done
for v in "3 2" "3 3" "3 6"; do set -- $v
  PSAM_TC_STAGES=$1 PSAM_BW_CTAS=$2 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_s$1_c$2.json 2> gpurun_out/r2c_bench_s$1_c$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_s$1_c$2.json"))
print("stages $1 ctas $2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernels_ms_per_step"])
PY
done
