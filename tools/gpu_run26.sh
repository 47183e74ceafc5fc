cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
b() { name=$1; tmo=$2; shift; shift
  timeout $tmo $T bench.py --gpus 8 "$@" > gpurun_out/r2x_$name.json 2> gpurun_out/r2x_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2x_$name.json") if l.startswith("{")][-1])
    ns=d.get("north_star_runs") or {}
    print("N=8 $name", round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), (d["parity_check"] or {}).get("ok"), {k:(round(v["value"]),round(v["ms_per_step"],3)) for k,v in ns.items()}, d["clocks"])
except Exception as e:
    print("$name FAILED", e)
    import subprocess; print(subprocess.run("grep -v '^frame\|^  File\|^    ' gpurun_out/r2x_$name.err | tail -c 1200", shell=True, capture_output=True, text=True).stdout)
PY
}
b p2p 170 --steps 400 --warmup 5 --no-cpu-baseline
b nccl 110 --steps 200 --warmup 5 --no-cpu-baseline --no-north-star-runs --p2p 0
