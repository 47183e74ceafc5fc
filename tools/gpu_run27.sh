cd $GRAFT_REPO_ROOT
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/symm_probe.py 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -30
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "most_conf or function_level" 2>&1 | tail -5
