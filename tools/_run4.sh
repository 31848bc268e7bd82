cd $GRAFT_REPO_ROOT
timeout -s USR1 -k 10 600 python -m pytest tests/test_gpu_oscr.py -x -q -m gpu > gpurun_out/r03_t_oscr.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r03_t_oscr.log
timeout -s USR1 -k 10 120 python tools/oscr_trace.py > gpurun_out/r03_oscr_trace.txt 2>&1; tail -32 gpurun_out/r03_oscr_trace.txt
for i in 1 2; do timeout -s USR1 -k 10 200 python bench.py --workload oscr > gpurun_out/r03_oscr_$i.json 2> gpurun_out/r03_oscr_$i.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r03_oscr_$i.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d.get('parity'))"; done
