"""Shared helpers for the tests."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_inf(a, b):
    """norm-wise relative difference max|a-b| / max|b|"""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


_hl = None


def hostlogic():
    global _hl
    if _hl is None:
        _hl = C.CDLL(os.path.join(ROOT, "tests", "_build", "libhostlogic.so"))
        _hl.hl_last_error.restype = C.c_char_p
    return _hl


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hl_solve(rowPtr, colPtr, dof, R, Val, ls, prec, faces, incL=None, res=None):
    """krylov.hpp (the product's solver control flow) driven by the TEST-ONLY serial host policy."""
    L = hostlogic()
    R = np.ascontiguousarray(R, np.float64).copy()
    Val = np.ascontiguousarray(Val, np.float64).copy()
    rowPtr = np.ascontiguousarray(rowPtr, np.int32); colPtr = np.ascontiguousarray(colPtr, np.int32)
    nF = len(faces)
    fn = np.array([len(f["nodes"]) for f in faces], np.int32)
    fd = np.array([f["dof"] for f in faces], np.int32)
    fb = np.array([f["bGrp"] for f in faces], np.int32)
    fg = np.concatenate([np.asarray(f["nodes"], np.int32) for f in faces]) if nF else np.zeros(0, np.int32)
    fv = np.concatenate([np.asarray(f["val"], np.float64).reshape(-1) for f in faces]) if nF else np.zeros(0)
    out = np.zeros(9)
    incL_a = None if incL is None else np.ascontiguousarray(incL, np.int32)
    res_a = None if res is None else np.ascontiguousarray(res, np.float64)
    ls = np.ascontiguousarray(ls, np.float64)
    rc = L.hl_solve(len(rowPtr) - 1, len(colPtr), _p(rowPtr), _p(colPtr), dof, _p(R), _p(Val), _p(ls), int(prec), nF,
                    _p(fn), _p(fd), _p(fb), _p(fg), _p(fv), _p(incL_a), _p(res_a), _p(out))
    if rc != 0:
        raise RuntimeError(L.hl_last_error().decode())
    keys = ["suc", "itr", "iNorm", "fNorm", "dB", "GM_itr", "CG_itr", "Resm", "Resc"]
    return R, Val, dict(zip(keys, out))


_eh = None


def elemhost():
    global _eh
    if _eh is None:
        _eh = C.CDLL(os.path.join(ROOT, "tests", "_build", "libelemhost.so"))
        _eh.host_fluid_assemble.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 7
        _eh.host_elem_tables.argtypes = [C.c_int, C.c_double] + [C.c_void_p] * 4
        _eh.host_face_integ.restype = C.c_double
        _eh.host_face_integ.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _eh.host_face_normals.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p]
        _eh.host_pk2cc.argtypes = [C.c_void_p] * 5
        _eh.host_pk2cc.restype = None
        _eh.host_bfolw_assemble.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 11
        _eh.host_bneu_assemble.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 11
    return _eh


def host_fluid_assemble(case):
    """fluid_elem.hpp (the product's Gauss-point arithmetic) run serially on the host by the TEST-ONLY harness
    tests/hostlogic/fluid_elem_host.cpp on a svfsiplus_b200.problem fluid case.  Returns R (nNo,4), Val (nnz,16)."""
    L = elemhost()
    m = case["mesh"]
    p = case["props"]
    tD = case["Ag"].shape[1]
    f = p.get("f", (0.0, 0.0, 0.0))
    par = np.array([p["dt"], p["am"], p["af"], p["gam"], p["rho"], f[0], f[1], f[2], p.get("Kinv", 0.0), p.get("viscType", 0),
                    p["mu"], p.get("mu_o", 0.0), p.get("lam", 0.0), p.get("a", 0.0), p.get("n", 0.0), tD,
                    int(p.get("mvMsh", False))], np.float64)
    ien = np.ascontiguousarray(m.ien, np.int32); x = np.ascontiguousarray(m.x, np.float64)
    Ag = np.ascontiguousarray(case["Ag"]); Yg = np.ascontiguousarray(case["Yg"]); Bf = np.ascontiguousarray(case["Bf"])
    rp = np.ascontiguousarray(case["rowPtr"], np.int32); cp = np.ascontiguousarray(case["colPtr"], np.int32)
    R = np.zeros((m.nNo, 4)); Val = np.zeros((len(cp), 16))
    rc = L.host_fluid_assemble(ien.shape[1], m.nEl, _p(ien), _p(x), None, _p(par), -1.0, _p(Ag), _p(Yg), _p(Bf), _p(rp), _p(cp),
                               _p(R), _p(Val))
    if rc != 0:
        raise RuntimeError(f"host_fluid_assemble: rc {rc}")
    return R, Val


def host_bneu_assemble(kind, mesh, IENb, gE, hg, Yg, rowPtr, colPtr, *, dt, af, gam, rho=0.0, bfs=0.0, mvMsh=False, Do=None):
    """face_elem.hpp (b_assem_neu_bc + gnnb + b_fluid / b_l_elas) run serially on the host by the TEST-ONLY harness."""
    L = elemhost()
    dof = 4 if kind == "fluid" else 3
    ien = np.ascontiguousarray(mesh.ien, np.int32); x = np.ascontiguousarray(mesh.x, np.float64)
    IENb = np.ascontiguousarray(IENb, np.int32); gE = np.ascontiguousarray(gE, np.int32)
    hg = np.ascontiguousarray(hg, np.float64); Yg = np.ascontiguousarray(Yg, np.float64)
    Do = None if Do is None else np.ascontiguousarray(Do, np.float64)
    rp = np.ascontiguousarray(rowPtr, np.int32); cp = np.ascontiguousarray(colPtr, np.int32)
    par = np.array([dt, af, gam, rho, bfs, Yg.shape[1], int(mvMsh)], np.float64)
    R = np.zeros((mesh.nNo, dof)); Val = np.zeros((len(cp), dof * dof))
    rc = L.host_bneu_assemble(0 if kind == "fluid" else 1, ien.shape[1], _p(ien), IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE),
                              _p(par), _p(x), _p(Do), _p(hg), _p(Yg), _p(rp), _p(cp), _p(R), _p(Val))
    if rc != 0:
        raise RuntimeError(f"host_bneu_assemble: rc {rc}")
    return R, Val


def host_face_integ(mesh, IENb, gE, s, l=0, u=None, geo=None, goff=0):
    """face_elem.hpp face_integ_terms summed serially on the host (TEST-ONLY harness).  s (nNo, nrows) or None (area)."""
    L = elemhost()
    ien = np.ascontiguousarray(mesh.ien, np.int32); x = np.ascontiguousarray(mesh.x, np.float64)
    IENb = np.ascontiguousarray(IENb, np.int32); gE = np.ascontiguousarray(gE, np.int32)
    sa = None if s is None else np.ascontiguousarray(s, np.float64)
    ga = None if geo is None else np.ascontiguousarray(geo, np.float64)
    u = l if u is None else u
    return L.host_face_integ(ien.shape[1], _p(ien), IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), _p(x), _p(ga),
                             0 if ga is None else ga.shape[1], goff, _p(sa), 1 if sa is None else sa.shape[1], l, u - l + 1)


def host_face_normals(mesh, IENb, gE, geo=None, goff=0):
    """face_elem.hpp face_normal_terms (fsi_ls_upd) accumulated on the host: sV (nNo, 3)."""
    L = elemhost()
    ien = np.ascontiguousarray(mesh.ien, np.int32); x = np.ascontiguousarray(mesh.x, np.float64)
    IENb = np.ascontiguousarray(IENb, np.int32); gE = np.ascontiguousarray(gE, np.int32)
    ga = None if geo is None else np.ascontiguousarray(geo, np.float64)
    sV = np.zeros((mesh.nNo, 3))
    rc = L.host_face_normals(ien.shape[1], _p(ien), IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), _p(x), _p(ga),
                             0 if ga is None else ga.shape[1], goff, _p(sV))
    if rc != 0:
        raise RuntimeError(f"host_face_normals: rc {rc}")
    return sV


def host_pk2cc(F, fl, *, iso, vol, C10=0.0, C01=0.0, Kpen=0.0, ho=None, Tfa=0.0, eta_s=0.0, kap=0.0):
    """solid_law.hpp pk2cc_iso on the host (TEST-ONLY harness): S (6: 00 11 22 01 12 20), Dm upper triangle (21)."""
    L = elemhost()
    ho = ho or {}
    keys = ("a", "b", "aff", "bff", "ass", "bss", "afs", "bfs", "khs")
    par = np.array([{"nHook": 0, "StVK": 1, "mStVK": 2, "HO": 3, "MR": 4, "HGO": 5, "Gucci": 6, "HO_ma": 7}[iso], {None: 0, "Quad": 1, "ST91": 2, "M94": 3}[vol], C10, C01, Kpen]
                   + [ho.get(k, 100.0 if k == "khs" else 0.0) for k in keys] + [Tfa, Tfa * eta_s, kap], np.float64)
    F = np.ascontiguousarray(F, np.float64); fl = np.ascontiguousarray(fl, np.float64)
    S6 = np.empty(6); Dm21 = np.empty(21)
    L.host_pk2cc(_p(par), _p(F), _p(fl), _p(S6), _p(Dm21))
    return S6, Dm21


def host_visc(model, mu, Nx, vx, F):
    """solid_law.hpp visc_point + visc_pair on the host (TEST-ONLY harness): Svis (3,3), Kvis_u, Kvis_v (a, b, 3, 3)."""
    import ctypes as C
    L = elemhost()
    L.host_visc.argtypes = [C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 6
    Nx = np.ascontiguousarray(Nx, np.float64); vx = np.ascontiguousarray(vx, np.float64); F = np.ascontiguousarray(F, np.float64)
    n = Nx.shape[0]
    S = np.zeros((3, 3)); Ku = np.zeros((n, n, 3, 3)); Kv = np.zeros((n, n, 3, 3))
    L.host_visc({"newt": 1, "pot": 2}[model], float(mu), n, _p(Nx), _p(vx), _p(F), _p(S), _p(Ku), _p(Kv))
    return S, Ku, Kv


def host_bfolw_assemble(mesh, IENb, gE, hg, Dg, rowPtr, colPtr, *, dt, af, beta=0.0, s=0, ustruct=False, am=1.0, gam=0.0):
    """face_elem.hpp face_follower_element (b_neu_folw_p: follower pressure on a struct / ustruct face) run serially on the host.
    Returns (R, Val) for struct, (R, Val, Kd) for ustruct."""
    L = elemhost()
    ien = np.ascontiguousarray(mesh.ien, np.int32); x = np.ascontiguousarray(mesh.x, np.float64)
    IENb = np.ascontiguousarray(IENb, np.int32); gE = np.ascontiguousarray(gE, np.int32)
    hg = np.ascontiguousarray(hg, np.float64); Dg = np.ascontiguousarray(Dg, np.float64)
    rp = np.ascontiguousarray(rowPtr, np.int32); cp = np.ascontiguousarray(colPtr, np.int32)
    afl = af * gam * dt if ustruct else af * beta * dt * dt
    afm = afl / am if ustruct else 0.0
    par = np.array([afl, afm, Dg.shape[1], s, float(ustruct)], np.float64)
    dof = 4 if ustruct else 3
    R = np.zeros((mesh.nNo, dof)); Val = np.zeros((len(cp), dof * dof)); Kd = np.zeros((len(cp), 12))
    rc = L.host_bfolw_assemble(ien.shape[1], _p(ien), IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), _p(par), _p(x), _p(Dg), _p(hg),
                               _p(rp), _p(cp), _p(R), _p(Val), _p(Kd))
    if rc != 0:
        raise RuntimeError(f"host_bfolw_assemble: rc {rc}")
    return (R, Val, Kd) if ustruct else (R, Val)
