#!/usr/bin/env python
"""Generator of tests/golden/p10_ns_counts.json: ONE Newton-iteration hot path of the compiled reference (oracle/_ref: the
reference's own construct_fluid + fsils_commuv + fsils_solve, LS NS with the pipe_RCR_3d <LS> block) on the benchmark-size mesh
P10 = pipe 96 x 96 x 181 = 10,008,576 TET4 as bench.py builds it (svfsiplus_b200.partition.local_slab_case: the pipe composed of 8
generation blocks, identical for 1, 2, 4, 8 ranks), on `--ranks` ranks of the in-process MPI stand-in (threads as ranks).

    python tests/golden/make_golden_p10.py [--dims 96 96 181] [--ranks 8] [--out tests/golden/p10_ns_counts.json]

Takes minutes and ~40 GB of host memory at P10; run offline in the build container (it needs /root/reference through
oracle/_ref).  The GPU test tests/test_gpu_parity.py::test_benchmark_size_matches_reference_golden asserts the counts
within +-1 and the norms to 1e-8 / the solution norms at the linear-solve tolerance.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _assemble(part):
    from oracle import refcase
    return refcase.reference_assemble(part)


def run(dims, ranks, ls="NS"):
    import multiprocessing as mp
    from oracle import ref, refcase
    from svfsiplus_b200 import partition as PT
    from svfsiplus_b200 import problem as P
    t0 = time.time()
    parts = [PT.local_slab_case(dims, r, ranks)[0] for r in range(ranks)]
    print(f"parts built in {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    with mp.get_context("fork").Pool(ranks) as pool:
        asm = pool.map(_assemble, parts)
    print(f"construct_fluid on {ranks} ranks: {time.time() - t0:.1f} s (slowest rank {max(a[4] for a in asm):.1f} s)", flush=True)
    rr = ref.RefRanks([dict(gnNo=p["gnNo"], gNodes=p["gNodes"], rowPtr=p["rowPtr"], colPtr=p["colPtr"],
                            faces=[dict(nodes=f["nodes"], dof=f["dof"], bGrp=f["bGrp"], val=f["val"]) for f in p["faces"]])
                       for p in parts])
    Rs = rr.commuv(4, [a[0] for a in asm])
    t0 = time.time()
    Xs, _, outs = rr.solve(4, refcase._ls_vector(P.LS_SETTINGS[ls]), ref.PREC_FSILS, Rs, [a[1] for a in asm],
                           parts[0]["incL"], parts[0]["res"])
    solve_s = time.time() - t0
    rr.close()
    print(f"fsils_solve: {solve_s:.1f} s", flush=True)
    gnNo = parts[0]["gnNo"]
    Xg = np.zeros((gnNo, 4)); Rg = np.zeros((gnNo, 4))
    for p, x, r in zip(parts, Xs, Rs):
        Xg[p["gNodes"]] = x; Rg[p["gNodes"]] = r
    o = outs[0]
    probe = np.linspace(0, gnNo - 1, 64).astype(np.int64)
    return {
        "dims": list(dims), "tets": int(6 * dims[0] * dims[1] * dims[2]), "gnNo": int(gnNo), "ls": ls, "reference_ranks": ranks,
        "itr": int(o["itr"]), "GM_itr": int(o["GM_itr"]), "CG_itr": int(o["CG_itr"]), "suc": bool(o["suc"]),
        "iNorm": float(o["iNorm"]), "fNorm": float(o["fNorm"]),
        "R_norm": [float(np.linalg.norm(Rg[:, j])) for j in range(4)],
        "X_norm": [float(np.linalg.norm(Xg[:, j])) for j in range(4)],
        "probe_nodes": [int(v) for v in probe], "X_probe": Xg[probe].tolist(), "R_probe": Rg[probe].tolist(),
        "reference_solve_s": solve_s, "reference_asm_s_slowest_rank": float(max(a[4] for a in asm)),
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[96, 96, 181])
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--ls", default="NS", choices=["NS", "GMRES"], help="<LS> block: NS (pipe_RCR_3d) or the plain GMRES variant of SURVEY 8(d)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.out is None:
        a.out = os.path.join(ROOT, "tests", "golden", f"p10_{a.ls.lower()}_counts.json")
    res = run(tuple(a.dims), a.ranks, a.ls)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print({k: res[k] for k in ("itr", "GM_itr", "CG_itr", "iNorm", "fNorm", "reference_solve_s")})
