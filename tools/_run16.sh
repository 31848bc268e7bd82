cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6; do timeout -s USR1 -k 10 300 python -m pytest tests/test_gpu_head.py -x -q -m gpu -k "fused or graph" 2>&1 | tail -1; done
timeout -s USR1 -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/r09_t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r09_t_all.log
