"""Multi-GPU parity (`-m gpu`, needs >= 2 devices): slab-partitioned pipe on 2 (and 4) GPUs, one process per
GPU over NCCL, against the compiled reference on one rank and on the same partition (threads as ranks)."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    from svfsiplus_b200 import backend as B
    return B.lib().b200_device_count()


def _run(world, dims, ls, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py"),
           *[str(d) for d in dims], ls]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("MULTIGPU_REPORT ")]
    assert lines, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(lines[-1][len("MULTIGPU_REPORT "):])


@pytest.mark.parametrize("world,ls", [(2, "NS"), (2, "GMRES"), (4, "NS")])
def test_partitioned_solve_matches_reference(world, ls):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    rep = _run(world, (8, 8, 16), ls, 29620 + world)
    assert all(rep["suc"]) and len(set(rep["itr"])) == 1
    assert max(rep["overlap_X"]) < 1e-9 * 1e6          # both owners of an overlap node hold the same solution
    if "X_vs_1rank" in rep:
        assert rep["R_vs_1rank"] < 1e-12
        assert rep["commu_R"] < 1e-14
        # the reference's own cross-partition tolerance (tests/conftest.py RTOL: velocity 1e-7 ... pressure 1e-6
        # after Newton convergence); a single 1e-3 linear solve is compared at the looser 1e-5 used elsewhere
        assert rep["X_vs_Nrank_ref"] < 1e-5 and rep["X_vs_1rank"] < 1e-4
        assert abs(rep["itr"][0] - rep["itr_Nrank_ref"]) <= 1
