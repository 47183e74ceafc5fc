cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_dist_nccl.py -m gpu -q -x 2>&1 | grep -v "^frame\|^  File\|^    " | tail -4
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-north-star-runs > gpurun_out/r2_final_2gpu.json 2> gpurun_out/r2_final_2gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_final_2gpu.json") if l.startswith("{")][-1])
print("N=2", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), (d["parity_check"] or {}).get("ok"), d["config"]["parallelism"][:90], d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
PY
