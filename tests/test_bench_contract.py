"""bench.py's reference arm and the keys both arms share (CPU; the GPU arm itself is exercised by the driver and by the GPU calls
recorded under profiles/)."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT
from conftest import needs_ref


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    return b


def test_config_is_the_same_object_in_both_arms():
    b = _bench_module()

    class A:
        dims = [96, 96, 181]
        ls = "NS"
    for n in (1, 2, 4, 8):
        c = b.make_config(A, n)
        assert c == b.make_config(A, n) and c["parallelism"] == f"dd{n}"
        assert "96x96x181 = 10008576 TET4" in c["workload"]          # every N solves the 10M-tet pipe (strong scaling headline)
    assert b.workload_dims(A, 8) == (96, 96, 181)


@needs_ref
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-dims", "8", "8", "16", "--ref-ranks", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["extrapolated"] is True
    assert d["run"]["sample_dims"] == [8, 8, 16] and d["run"]["extrapolated"] is True
    b = _bench_module()

    class A:
        dims = [96, 96, 181]
        ls = "NS"
    assert d["config"] == b.make_config(A, 1)
    assert d["metric"] == b.METRIC and d["unit"] == b.UNIT


@needs_ref
def test_reference_arm_under_torchrun_only_rank_zero_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--ref-dims", "8", "8", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_other_workloads_answer_the_reference_arm_with_unavailable():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "struct_block"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0
    d = json.loads(r.stdout.strip())
    assert d["impl"] == "reference" and "unavailable" in d
