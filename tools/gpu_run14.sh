set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -k "graphed or cca or components" 2>&1 | tail -3
b() { name=$1; shift
  env X=1 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2k_$name.json"))
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
b base
b split --split-streams 1
b split_l6 --split-streams 1 --lanes 6
b split_l3 --split-streams 1 --lanes 3
b split_l2 --split-streams 1 --lanes 2
export PSAM_BW_CTAS=2
b split_c2 --split-streams 1
unset PSAM_BW_CTAS
timeout 300 python tools/trace_timeline.py run --steps 12 --lanes 4 --split 1
timeout 300 python tools/trace_timeline.py show gpurun_out/trace.npy > gpurun_out/trace_split.txt; head -32 gpurun_out/trace_split.txt
