set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
b() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2j_$name.json"))
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), d["roofline"]["kernels_ms_per_step"])
except Exception as e:
    print("$name FAILED", e)
PY
}
b base X=1
b c2 PSAM_BW_CTAS=2
b c3 PSAM_BW_CTAS=3
b c5 PSAM_BW_CTAS=5
timeout 300 python tools/trace_timeline.py run --steps 12 --lanes 4
timeout 300 python tools/trace_timeline.py show gpurun_out/trace.npy | head -40
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs --split-streams 1 > gpurun_out/r2j_split.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2j_split.json')); print('split', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs --lanes 6 > gpurun_out/r2j_l6.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2j_l6.json')); print('lanes6', d['value'], d['ms_per_step'])"
