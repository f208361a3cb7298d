"""CPU checks of oracle/_ref/svmultiphysics_b200 (the reference's own main() with the B200 backend registered by
oracle/patch_reference.py): registering the backend must not change anything for a case that does not
select it, and selecting it without a GPU must fail loudly - there is no CPU fallback.  (The GPU side is tests/test_reference_main.py.)"""
import os
import subprocess

import numpy as np
import pytest

from util import ROOT
from svfsiplus_b200 import sv_io as IO

EXE_REF = os.path.join(ROOT, "oracle", "_ref", "svmultiphysics_ref")
EXE_B200 = os.path.join(ROOT, "oracle", "_ref", "svmultiphysics_b200")
needs = pytest.mark.skipif(not (os.path.exists(EXE_REF) and os.path.exists(EXE_B200)),
                           reason="oracle/_ref/svmultiphysics_{ref,b200} not built (needs the reference sources)")


def _export():
    import importlib.util
    spec = importlib.util.spec_from_file_location("export_case", os.path.join(ROOT, "tools", "export_case.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    return ex


@needs
@pytest.mark.parametrize("kind", ["pipe", "block", "ustruct"])
def test_registered_backend_leaves_fsils_runs_bit_identical(tmp_path, kind):
    ex = _export()
    outs = []
    for name, exe in (("ref", EXE_REF), ("b200", EXE_B200)):
        d = tmp_path / name
        if kind == "pipe":
            ex.export_pipe(str(d), (4, 4, 6), steps=2)
        elif kind == "block":
            ex.export_block(str(d), 3, "hex", steps=2)
        else:
            ex.export_block(str(d), 3, "tet", steps=1, phys="ustruct")
        r = subprocess.run([exe, "solver.xml"], cwd=d, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(d / "1-procs")
    last = sorted(f for f in os.listdir(outs[0]) if f.startswith("result_"))[-1]
    a, b = IO.read_vtk(outs[0] / last), IO.read_vtk(outs[1] / last)
    assert list(a["point_data"]) == list(b["point_data"])
    for k in a["point_data"]:
        assert np.array_equal(a["point_data"][k], b["point_data"][k]), k
    # the history differs only in its wall-clock columns (time stamp, % of the time spent in the linear solver)
    strip = lambda p: [l.split("[", 1)[1].rsplit(None, 1)[0] for l in open(p).read().splitlines() if "[" in l]
    assert strip(outs[0] / "histor.dat") == strip(outs[1] / "histor.dat")


@needs
@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present: the backend would run")
def test_selecting_the_backend_without_a_gpu_fails_loudly(tmp_path):
    ex = _export()
    ex.export_pipe(str(tmp_path / "c"), (4, 4, 6), steps=1, linear_algebra="b200")
    r = subprocess.run([EXE_B200, "solver.xml"], cwd=tmp_path / "c", capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CUDA device available: this backend has no CPU fallback" in (r.stdout + r.stderr)


@needs
def test_unpatched_reference_rejects_the_backend_name(tmp_path):
    """The same solver.xml through the UNPATCHED reference: `b200` is not a LinearAlgebra type there (Parameters.cpp:2405-2411)."""
    ex = _export()
    ex.export_pipe(str(tmp_path / "c"), (4, 4, 6), steps=1, linear_algebra="b200")
    r = subprocess.run([EXE_REF, "solver.xml"], cwd=tmp_path / "c", capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "Unknown TYPE 'b200'" in (r.stdout + r.stderr)


def test_patch_script_asserts_its_anchors(tmp_path):
    """oracle/patch_reference.py fails instead of silently producing an unpatched file when an anchor is missing."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("patch_reference", os.path.join(ROOT, "oracle", "patch_reference.py"))
    pr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pr)
    with pytest.raises(AssertionError):
        pr.patch_main('#include "x.h"\nint main() { return 0; }\n')
    with pytest.raises((AssertionError, ValueError)):
        pr.patch_consts("enum class Other { a };")
    ok = pr.patch_consts("enum class LinearAlgebraType {\n  none,\n  fsils,\n  petsc,\n  trilinos\n};\n")
    assert "b200" in ok and ok.count("trilinos") == 1
