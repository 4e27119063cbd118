"""bench.py's reference arm (the CPU restatement timed on the host cores) runs without a GPU: its JSON line must
carry the keys the driver reads, with the unit and direction of our own arm.  (Our own arm needs a B200; its line
is checked against the same key list by the committed profile of the last run.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _line(args, env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _line(["--impl", "reference", "--workload", "config2", "--steps", "1", "--warmup", "0"], {"NRT_REF_STEP_SECONDS": "0.5"})
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "Mrays/sec" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "linear resolution" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_bench_line_has_the_contract_keys():
    # the last line our own arm printed on a B200 (profiles/): the keys of the contract + roofline / cpu_baseline / clocks
    path = os.path.join(ROOT, "profiles", "r02b_bench_config4_n1.json")
    d = next(json.loads(l) for l in open(path) if l.startswith("{"))
    assert BASE_KEYS <= set(d) and {"roofline", "cpu_baseline", "clocks", "parity"} <= set(d)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["parity"]["matches_oracle"] is True and d["clocks"]["reasons"] == []
