cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_dist_nccl.py -m gpu -q -x 2>&1 | tail -15
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
b() { name=$1; shift
  timeout 200 $T bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs "$@" > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2w_$name.json") if l.startswith("{")][-1])
    print("N=2 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), (d["parity_check"] or {}).get("ok"))
except Exception as e:
    print("$name FAILED", e)
    import subprocess; print(subprocess.run("tail -c 1500 gpurun_out/r2w_$name.err", shell=True, capture_output=True, text=True).stdout)
PY
}
b p2p --p2p 1
b nccl --p2p 0
