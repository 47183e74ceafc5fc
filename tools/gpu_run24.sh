cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
b() { name=$1; shift
  timeout 200 $T bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2t_$name.json") if l.startswith("{")][-1])
    print("N=2 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["parity_check"])
except Exception as e:
    print("$name FAILED", e)
PY
}
b base
PSAM_TC_RESERVE_SMS=4 b res4
PSAM_TC_RESERVE_SMS=8 b res8
b l6 --lanes 6
b ch2 --nccl-channels 2
PSAM_TC_RESERVE_SMS=4 b res4_l6 --lanes 6
b gc --graph-collectives 1
timeout 300 python -m pytest tests/test_dist_nccl.py -m gpu -q 2>&1 | tail -3
