#!/bin/bash
# Round-2 ncu evidence (one B200): the launch list of a short bench run and one --set full capture of each big kernel.
#   gpurun --timeout 1500 -- 'bash tools/profile_round2.sh'   ->  gpurun_out/r2_*   (summaries are copied to profiles/ here)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-north-star-runs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv $B > gpurun_out/r2_launches_final.log 2>&1
for k in k_match_ts k_blocks_warp k_components; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r2_final_$k $B --lanes 1 --no-graphs > gpurun_out/r2_final_$k.log 2>&1
done
ls -la gpurun_out | tail -12
