cd $GRAFT_REPO_ROOT
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-north-star-runs --algo 3 --lanes 1 --no-graphs"
PSAM_TC_CW=8 PSAM_TC_PF=2 PSAM_TC_STAGES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_match_tc -s 4 -c 1 -f -o gpurun_out/r2_fused_b $B > gpurun_out/r2_fused_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
