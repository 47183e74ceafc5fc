set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
for c in 2 6; do
PSAM_BW_CTAS=$c timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blocks_warp -s 4 -c 1 -o gpurun_out/r2_prof_bw_c$c -f python bench.py --steps 3 --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline > gpurun_out/r2_prof_bw_c$c.log 2>&1
done
