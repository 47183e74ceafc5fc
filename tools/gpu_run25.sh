cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or golden or graphed" 2>&1 | tail -3
b() { name=$1; shift
  timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2u_$name.json"))
    k=d["roofline"]["kernels_ms_per_step"]
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {a:b for a,b in k.items() if 'match' in a or 'pack_q' in a})
except Exception as e:
    print("$name FAILED", e)
PY
}
b auto
PSAM_TC_FULL_GRID=1 b full
b auto2
b split --split-streams 1
b l5 --lanes 5
b l3 --lanes 3
