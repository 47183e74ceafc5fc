set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/trace_timeline.py run --steps 12 --lanes 4
timeout 300 python tools/trace_timeline.py show gpurun_out/trace.npy
cp gpurun_out/trace.npy gpurun_out/trace_c3.npy
PSAM_BW_CTAS=2 timeout 300 python tools/trace_timeline.py run --steps 12 --lanes 4
cp gpurun_out/trace.npy gpurun_out/trace_c2.npy
timeout 300 python tools/trace_timeline.py show gpurun_out/trace_c2.npy
