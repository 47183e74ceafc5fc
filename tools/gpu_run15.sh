set -x
cd $GRAFT_REPO_ROOT
b() { name=$1; shift
  timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2l_$name.json"))
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
export CUDA_DEVICE_MAX_CONNECTIONS=32
b mc32_base
b mc32_split --split-streams 1
b mc32_split_l6 --split-streams 1 --lanes 6
b mc32_l6 --lanes 6
b mc32_split_l8 --split-streams 1 --lanes 8
timeout 300 python tools/trace_timeline.py run --steps 12 --lanes 4 --split 1
timeout 300 python tools/trace_timeline.py show gpurun_out/trace.npy > gpurun_out/trace_split_mc32.txt; sed -n 9,60p gpurun_out/trace_split_mc32.txt
export CUDA_DEVICE_MAX_CONNECTIONS=1
b mc1_base
