cd $GRAFT_REPO_ROOT
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-north-star-runs --lanes 1 --no-graphs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_match_ts -s 4 -c 1 -f -o gpurun_out/r2_ts $B > gpurun_out/r2_ts.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
