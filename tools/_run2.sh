cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_head.py -x -q -m gpu -k "graph or handover or headline" > gpurun_out/r03_t_head2.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r03_t_head2.log
for rep in 1 2 3; do
for v in default noflags serial; do
  case $v in
    default) E="";;
    noflags) E="OSB_NO_ROW_FLAGS=1";;
    serial) E="OSB_SERIAL_BACKWARD=1";;
  esac
  for st in 20 300; do
    env $E timeout 300 python bench.py --steps $st --warmup 5 --no-oscr --no-sub --no-e2e > gpurun_out/r03b_${v}_${st}_$rep.json 2> gpurun_out/r03b_err.txt || echo "FAIL $v $st"
  done
done
done
python - <<'PY'
import json,glob,statistics
for v in ["default","noflags","serial"]:
    for st in (20,300):
        xs=[]
        for f in sorted(glob.glob(f"gpurun_out/r03b_{v}_{st}_*.json")):
            try: xs.append(round(json.loads(open(f).read().strip().splitlines()[-1])["ms_per_step"]*1e3,1))
            except Exception as e: xs.append(None)
        print(v, st, xs)
PY
