cd $GRAFT_REPO_ROOT
timeout -s USR1 -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/r09_t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r09_t_all.log
timeout -s USR1 -k 10 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r09_bench_default20.json 2> gpurun_out/r09_bench_default20.err; echo "bench rc=$?"
timeout 120 python tools/head_once.py > gpurun_out/r09_head_once.log 2>&1; echo "head_once rc=$?"
OSB_NO_COOP_LAUNCH=1 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:head_|dw_reduce|label_stats' -s 6 -c 3 -f -o gpurun_out/r09_head_full python tools/head_once.py > gpurun_out/r09_ncu_head.log 2>&1; echo "ncu full rc=$?"
OSB_NO_COOP_LAUNCH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r09_launches.csv python bench.py --steps 2 --warmup 3 --no-sub --no-oscr --no-e2e > gpurun_out/r09_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r09_bench_default20.json").read().strip().splitlines()[-1])
print("step us", d["ms_per_step"]*1e3, "e2e", d["e2e"]["value"], "launch", d["config"]["launch"])
print(json.dumps(d["roofline"])[:1200])
print("oscr", d["oscr"]["value"], d["oscr"]["roofline"]["frac"], "bg", d["bg"]["us_per_step"], "resnet", d["resnet"]["value"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
