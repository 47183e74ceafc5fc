set -x
cd $GRAFT_REPO_ROOT
b() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2i_$name.json"))
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4))
except Exception as e:
    print("$name FAILED", e)
PY
}
b base X=1
b nocomp PSAM_EXPERIMENT_SKIP_COMPONENTS=1
b noblocks PSAM_EXPERIMENT_SKIP_BLOCKS=1
b neither PSAM_EXPERIMENT_SKIP_COMPONENTS=1 PSAM_EXPERIMENT_SKIP_BLOCKS=1
b nocomp_c2 PSAM_EXPERIMENT_SKIP_COMPONENTS=1 PSAM_BW_CTAS=2
b nocomp_c6 PSAM_EXPERIMENT_SKIP_COMPONENTS=1 PSAM_BW_CTAS=6
b s4_neither PSAM_EXPERIMENT_SKIP_COMPONENTS=1 PSAM_EXPERIMENT_SKIP_BLOCKS=1 PSAM_TC_STAGES=4
