cd $GRAFT_REPO_ROOT
b() { name=$1; shift
  timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2r_$name.json 2> gpurun_out/r2r_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2r_$name.json"))
    k=d["roofline"]["kernels_ms_per_step"]
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {a:b for a,b in k.items() if 'match' in a or 'pack_q' in a})
except Exception as e:
    print("$name FAILED", e)
PY
}
for nch in 224 208 192; do for st in 3 4 5; do
PSAM_TS_NCH=$nch PSAM_TS_STAGES=$st b ts_n${nch}_s${st}
done; done
b ts_split --split-streams 1
b ts_l6 --lanes 6
b ts_l3 --lanes 3
b packed --algo 2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
