set -x
cd $GRAFT_REPO_ROOT
run() { # name, extra args
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-north-star-runs "$@" > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_$name.json").read().strip().splitlines()[-1])
    print("N=2 $name", round(d["value"]), round(d["ms_per_step"],4), "host", round(d["host_enqueue_ms_per_step"],3), d["parity_check"]["ok"])
except Exception as e:
    print("N=2 $name FAILED", e)
PY
}
run base
run ch4 --nccl-channels 4
run ch8 --nccl-channels 8
run gc --graph-collectives 1
run gc_ch4 --graph-collectives 1 --nccl-channels 4
run split --split-streams 1
tail -5 gpurun_out/r2h_gc.err
for s in 0 1; do
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs --split-streams $s > gpurun_out/r2h_n1_split$s.json 2> gpurun_out/r2h_n1_split$s.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2h_n1_split$s.json"))
print("N=1 split $s", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
for l in 6 8; do
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs --lanes $l > gpurun_out/r2h_n1_l$l.json 2> gpurun_out/r2h_n1_l$l.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2h_n1_l$l.json"))
print("N=1 lanes $l", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
