cd $GRAFT_REPO_ROOT
timeout 500 python -m pytest tests/test_dist_nccl.py -m gpu -q -x 2>&1 | grep -v "^frame\|^  File\|^    " | tail -15
