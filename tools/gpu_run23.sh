cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
b() { name=$1; shift
  timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s_$name.json"))
    k=d["roofline"]["kernels_ms_per_step"]
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {a:b for a,b in k.items() if 'match' in a or 'pack_q' in a})
except Exception as e:
    print("$name FAILED", e)
PY
}
b ts
b ts_split --split-streams 1
b ts_l6 --lanes 6
PSAM_TS_STAGES=3 b ts_s3
b packed --algo 2
