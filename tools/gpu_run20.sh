cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tensor_core or fused or match" 2>&1 | tail -4
b() { name=$1; shift
  timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2p_$name.json"))
    k=d["roofline"]["kernels_ms_per_step"]
    print("N=1 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {a:b for a,b in k.items() if 'match' in a or 'pack_q' in a})
except Exception as e:
    print("$name FAILED", e)
PY
}
b base
for cfg in "8 4" "8 3" "8 2" "4 3" "4 2" "4 1"; do set -- $cfg; for st in 3 4; do
PSAM_TC_CW=$1 PSAM_TC_PF=$2 PSAM_TC_STAGES=$st b fused_cw$1_pf$2_st${st} --algo 3
done; done
