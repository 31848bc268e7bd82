#!/bin/bash
# Developer tool: the default bench line at N GPUs exactly as the driver launches it (--steps 20 --warmup 5), with a
# Python stack dump on SIGUSR1 if it gets stuck.   gpurun --gpus N -- 'bash tools/bench_full_n.sh N'
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}
mkdir -p gpurun_out
timeout -s USR1 -k 20 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus "$N" --steps 20 --warmup 5 > "gpurun_out/bench_n${N}_full.json" 2> "gpurun_out/bench_n${N}_full.err"
echo "bench rc=$?"
tail -5 "gpurun_out/bench_n${N}_full.err"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n${N}_full.json").read().strip().splitlines()[-1])
print("step us", d["ms_per_step"] * 1e3, "value", d["value"], "e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print("parity", json.dumps(d["parity"])[:900])
o = d.get("oscr"); print("oscr", o and (o["value"], o["config"]["launch"], o.get("parity")))
b = d.get("bg"); print("bg", b and (b["us_per_step"], b["value"], b["parity"].get("sharded")))
PY
