"""The reference arm of bench.py (`--impl reference`) runs on host cores only, so its contract can be checked here:
one JSON line on stdout with the keys the driver reads, the reference's own files behind it (`kind: "reference"`),
and nothing at all from ranks other than 0 under a multi-process launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "bg", "--steps", "3",
                           "--warmup", "3", *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 3 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("garbage open-set head fwd+bwd, 2048->152 classes, 1024 bf16 feature rows per GPU")
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and "1024 rows per step" in cb["sample"]
    from oracle import ref_loader
    if ref_loader.available():                       # /root/reference mounted, or its files vendored under oracle/_ref
        assert cb["kind"] == "reference"


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
