set -x
cd $GRAFT_REPO_ROOT
for k in k_blocks_warp k_components; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/r2_prof_$k -f python bench.py --steps 3 --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline > gpurun_out/r2_prof_$k.log 2>&1
done
ls -la gpurun_out/
