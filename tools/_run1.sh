set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_head.py -x -q -m gpu > gpurun_out/r03_t_head.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r03_t_head.log
for v in default noflags serial; do
  case $v in
    default) E="";;
    noflags) E="OSB_NO_ROW_FLAGS=1";;
    serial) E="OSB_SERIAL_BACKWARD=1";;
  esac
  env $E timeout 300 python bench.py --steps 300 --warmup 10 --no-oscr --no-sub --no-e2e > gpurun_out/r03_bench_$v.json 2> gpurun_out/r03_bench_$v.err; echo "bench $v rc=$?"
  env $E timeout 300 python bench.py --steps 20 --warmup 5 --no-oscr --no-sub --no-e2e > gpurun_out/r03_bench20_$v.json 2>> gpurun_out/r03_bench_$v.err; echo "bench20 $v rc=$?"
done
python - <<'PY'
import json
for v in ["default","noflags","serial"]:
    for pre in ["r03_bench_","r03_bench20_"]:
        try:
            d=json.loads(open(f"gpurun_out/{pre}{v}.json").read().strip().splitlines()[-1])
            print(pre+v, round(d["ms_per_step"]*1e3,1), d["roofline"].get("kernels_us"), d["parity"]["head"].get("dw_rel"), d["clocks"])
        except Exception as e:
            print(pre+v, "ERR", e)
PY
