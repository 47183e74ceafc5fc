set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_nccl.py tests/test_gpu_parity.py -m gpu -q -k "nccl or graphed or compact" 2>&1 | tail -8
for lanes in 4 6; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --lanes $lanes > gpurun_out/r2g_bench2_l$lanes.json 2> gpurun_out/r2g_bench2_l$lanes.err
tail -c 1500 gpurun_out/r2g_bench2_l$lanes.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_bench2_l$lanes.json").read().strip().splitlines()[-1])
print("N=2 lanes $lanes", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_check"], {k:(v["value"],v["ms_per_step"]) for k,v in (d["north_star_runs"] or {}).items()})
PY
done
