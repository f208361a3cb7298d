"""Golden fixtures from the COMPLETE reference solver (oracle/_ref/svmultiphysics_ref: the reference's own main() on the VTK-free
replacements, one rank) run on case directories written by tools/export_case.py:
    python tests/golden/make_golden_full_reference.py
Writes tests/golden/full_reference_runs.npz: the nodal result fields of the last saved step and the restart state, per case.  They are
what an end-to-end run of the product on the same exported case has to reproduce within the reference harness's tolerances
(svfsiplus_b200.sv_io.compare_results); tests/test_sv_io.py checks that the reference still produces them bit for bit."""
import importlib.util
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from svfsiplus_b200 import sv_io as IO  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(ROOT, "oracle", "_ref", "svmultiphysics_ref")

CASES = {"pipe_4_4_6": ("pipe", (4, 4, 6)), "block_hex_3": ("block", (3, "hex")), "block_tet_3": ("block", (3, "tet"))}


def run_case(name, workdir):
    spec = importlib.util.spec_from_file_location("export_case", os.path.join(ROOT, "tools", "export_case.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    kind, arg = CASES[name]
    out = os.path.join(workdir, name)
    if kind == "pipe":
        ex.export_pipe(out, arg, steps=2)
    else:
        ex.export_block(out, arg[0], arg[1], steps=2)
    subprocess.run([EXE, "solver.xml"], cwd=out, check=True, capture_output=True, timeout=600)
    vt = IO.read_vtk(os.path.join(out, "1-procs", "result_002.vtu"))
    return {f"{name}/{k}": v for k, v in vt["point_data"].items()}


def main():
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name in CASES:
            out.update(run_case(name, d))
    np.savez_compressed(os.path.join(HERE, "full_reference_runs.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
