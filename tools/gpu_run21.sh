cd $GRAFT_REPO_ROOT
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core" 2>&1 | tail -15
b() { name=$1; shift
  timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2q_$name.json"))
    k=d["roofline"]["kernels_ms_per_step"]
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {a:b for a,b in k.items() if 'match' in a or 'pack_q' in a})
except Exception as e:
    print("$name FAILED", e)
PY
}
b ts_st3 --algo 3
PSAM_TC_STAGES=4 b ts_st4 --algo 3
b base
