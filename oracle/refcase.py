"""TEST INFRASTRUCTURE ONLY (oracle).  Runs a svfsiplus_b200.problem case through the compiled
reference (oracle/_ref/libsvref.so): construct_fluid for R/Val, then fsils_solve."""
from __future__ import annotations

import numpy as np

from . import ref


def _ls_vector(ls):
    ls_type, RI, GM, CG = ls
    GM = GM or (1e-2, 1e-10, 2, 100)
    CG = CG or (0.2, 1e-10, 500, 0)
    return ref.ls_params(ls_type, RI[0], RI[1], RI[2], RI[3], gm=(GM[0], GM[1], GM[2], GM[3]), cg=(CG[0], CG[1], CG[2]))


def reference_assemble(case):
    m = case["mesh"]
    ra = ref.RefAssembly(m.x, m.ien)
    p = dict(case["props"])
    visc = None
    if p.get("viscType", 0) != 0:
        visc = [p["viscType"], p["mu"], p.get("mu_o", 0.0), p.get("lam", 0.0), p.get("a", 0.0), p.get("n", 0.0)]
    R, Val, secs = ra.fluid(case["Ag"], case["Yg"], case["Bf"], dt=p["dt"], am=p["am"], af=p["af"], gam=p["gam"],
                            rho=p["rho"], mu=p["mu"], f=p.get("f", (0.0, 0.0, 0.0)), Kinv=p.get("Kinv", 0.0),
                            visc=visc, mvMsh=p.get("mvMsh", False))
    rowPtr, colPtr = ra.csr()
    ra.close()
    return R, Val, rowPtr, colPtr, secs


def reference_solve(case, R, Val, ls, prec=ref.PREC_FSILS):
    m = case["mesh"]
    part = dict(gnNo=m.nNo, gNodes=np.arange(m.nNo), rowPtr=case["rowPtr"], colPtr=case["colPtr"],
                faces=[dict(nodes=f["nodes"], dof=f["dof"], bGrp=f["bGrp"], val=f["val"]) for f in case["faces"]])
    rr = ref.RefRanks([part])
    X, Vs, out = rr.solve(R.shape[1], _ls_vector(ls), prec, [R], [Val], case["incL"], case["res"])
    rr.close()
    return X[0], out[0]


def reference_step(case, ls="NS", prec=ref.PREC_FSILS):
    from svfsiplus_b200.problem import LS_SETTINGS
    ls = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    R, Val, rowPtr, colPtr, _ = reference_assemble(case)
    X, out = reference_solve(case, R, Val, ls, prec)
    return R, Val, X, out


def reference_assemble_solid(case):
    """construct_dsolid / construct_l_elas on a svfsiplus_b200.problem.block_case."""
    m = case["mesh"]
    ra = ref.RefAssembly(m.x, m.ien)
    p = dict(case["props"])
    ra.set_fibers(case.get("fN"))
    R, Val, secs = ra.solid(case["kind"], case["Ag"], case["Yg"], case["Dg"], case["Bf"], Do=case.get("Do"),
                            pS0=case.get("pS0"), pstEq=case.get("pstEq", False), **p)
    rowPtr, colPtr = ra.csr()
    tabs = ra.tables()
    case["_ref_pSn"], case["_ref_pSa"] = ra.pSn, ra.pSa          # com_mod.pSn / pSa when the case has pstEq
    ra.close()
    return R, Val, rowPtr, colPtr, secs, tabs


def reference_solid_step(case, ls, prec=ref.PREC_FSILS):
    from svfsiplus_b200.problem import LS_SETTINGS
    ls = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    R, Val, rowPtr, colPtr, _, _ = reference_assemble_solid(case)
    X, out = reference_solve(case, R, Val, ls, prec)
    return R, Val, X, out


def reference_assemble_fsi(case):
    m = case["mesh"]
    ra = ref.RefAssembly(m.x, m.ien)
    t = case["time"]
    R, Val, secs = ra.fsi(case["elem_dmn"], case["Ag"], case["Yg"], case["Dg"], case["Bf"], dt=t["dt"], am=t["am"], af=t["af"],
                          gam=t["gam"], beta=t["beta"], fluid=case["fluid"], solid=case["solid"], pS0=case.get("pS0"))
    ra.close()
    return R, Val, secs


def reference_fsi_step(case, ls, prec=ref.PREC_FSILS):
    from svfsiplus_b200.problem import LS_SETTINGS
    ls = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    R, Val, _ = reference_assemble_fsi(case)
    X, out = reference_solve(case, R, Val, ls, prec)
    return R, Val, X, out


def reference_assemble_ustruct(case, with_r=False):
    m = case["mesh"]
    ra = ref.RefAssembly(m.x, m.ien)
    ra.set_fibers(case.get("fN"))
    R, Val, Kd, secs = ra.ustruct(case["Ag"], case["Yg"], case["Dg"], case["Bf"], Ad=case["Ad"] if with_r else None, **case["props"])
    ra.close()
    return R, Val, Kd, secs


def reference_ustruct_step(case, ls, with_r=True, prec=ref.PREC_FSILS):
    from svfsiplus_b200.problem import LS_SETTINGS
    ls = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    R, Val, Kd, _ = reference_assemble_ustruct(case, with_r=with_r)
    X, out = reference_solve(case, R, Val, ls, prec)
    return R, Val, Kd, X, out


def reference_step_ranks(case, ls, nranks):
    """One Newton-iteration hot path of the reference on `nranks` ranks of the in-process MPI stand-in (threads as
    ranks, oracle/mpi_stub): the case is cut into z-slabs like a partitioned run, every rank assembles its own
    elements (construct_fluid; no communication), then commu(R) + fsils_solve run on all ranks together.
    Returns dict(asm_s = slowest rank's construct_fluid, solve_s, wall_s, itr, GM_itr, CG_itr, nranks)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    from svfsiplus_b200 import partition as PT
    from svfsiplus_b200.problem import LS_SETTINGS
    ls = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    if nranks <= 1:
        t0 = time.perf_counter()
        R, Val, _, _, t_asm = reference_assemble(case)
        X, out = reference_solve(case, R, Val, ls)
        return dict(asm_s=t_asm, solve_s=float(out["wall_s"]), wall_s=time.perf_counter() - t0, itr=int(out["itr"]),
                    GM_itr=int(out["GM_itr"]), CG_itr=int(out["CG_itr"]), nranks=1)
    parts = PT.split_case(case, nranks)
    # Assembly: one PROCESS per rank (like MPI ranks).  Threads would share the reference's static allocation
    # counters (Array<T>::num_allocated) and glibc arenas, which serialises its many per-element heap allocations.
    # The time of a rank is the time inside its own construct_fluid; the slowest rank counts.
    try:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(nranks) as pool:
            asm = pool.map(reference_assemble, parts)
    except Exception:
        with ThreadPoolExecutor(nranks) as ex:
            asm = list(ex.map(reference_assemble, parts))
    t0, t1 = 0.0, max(a[4] for a in asm)
    rr = ref.RefRanks([dict(gnNo=p["gnNo"], gNodes=p["gNodes"], rowPtr=p["rowPtr"], colPtr=p["colPtr"],
                            faces=[dict(nodes=f["nodes"], dof=f["dof"], bGrp=f["bGrp"], val=f["val"]) for f in p["faces"]])
                       for p in parts])
    t2 = time.perf_counter()
    Rs = rr.commuv(4, [a[0] for a in asm])
    Xs, _, outs = rr.solve(4, _ls_vector(ls), ref.PREC_FSILS, Rs, [a[1] for a in asm], case["incL"], case["res"])
    t3 = time.perf_counter()
    rr.close()
    # lhs_create / bc_create (t1..t2) is one-time set-up in the reference, not part of a Newton iteration
    return dict(asm_s=t1 - t0, solve_s=t3 - t2, wall_s=(t1 - t0) + (t3 - t2), itr=int(outs[0]["itr"]),
                GM_itr=int(outs[0]["GM_itr"]), CG_itr=int(outs[0]["CG_itr"]), nranks=nranks)
