"""TEST INFRASTRUCTURE ONLY (oracle).  ctypes front-end of oracle/_ref/libsvref.so — the reference's
own sources (Code/Source/liner_solver/*.cpp, Code/Source/solver/{fluid,lhsa,nn,fs,...}.cpp) compiled
unmodified by oracle/Makefile, driven through oracle/ref_harness.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  Nothing under svfsiplus_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVREF_LIB selects another build of the same sources (bench.py's -O3 timing leg: oracle/_ref/o3/libsvref.so); the parity tests
# always use the default -O2 build
LIB_PATH = os.environ.get("SVREF_LIB") or os.path.join(_HERE, "_ref", "libsvref.so")
O3_LIB_PATH = os.path.join(_HERE, "_ref", "o3", "libsvref.so")

LS_CG, LS_GMRES, LS_NS, LS_BICGS = 798, 797, 796, 795      # L/fils_struct.hpp:70-76
PREC_FSILS, PREC_RCS = 701, 709                              # S/consts.h:426

DROPIN_PATH = os.path.join(_HERE, "_ref", "libsvdropin.so")

_lib = None
_dropin = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def dropin_available() -> bool:
    return os.path.exists(DROPIN_PATH)


def dropin_lib():
    """The reference objects + the B200LinearAlgebra plug-in class linked against libsvb200.so."""
    global _dropin
    if _dropin is None:
        if not dropin_available():
            raise RuntimeError(f"{DROPIN_PATH} missing: run `make -C oracle ref` after building libsvb200.so")
        L = C.CDLL(DROPIN_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_asm_create.restype = C.c_void_p
        L.ref_asm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.ref_asm_destroy.argtypes = [C.c_void_p]
        L.ref_dropin_fluid_step.argtypes = ([C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 5 + [C.c_void_p, C.c_double]
                                            + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 8)
        L.ref_dropin_fluid_face_step.argtypes = ([C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 6 + [C.c_void_p] * 4 + [C.c_int]
                                                 + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 8)
        L.ref_dropin_solid_step.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int]
                                            + [C.c_void_p] * 8)
        _dropin = L
    return _dropin


def dropin_solid_step(case, ls, mode):
    """Like dropin_fluid_step for a svfsiplus_b200.problem.block_case (struct / lelas / mesh, dof 3)."""
    L = dropin_lib()
    m = case["mesh"]
    x = _c(m.x, np.float64); ien = _c(m.ien, np.int32)
    h = L.ref_asm_create(m.nNo, m.nEl, ien.shape[1], _p(ien), _p(x), 1, -1.0)
    if not h:
        raise RuntimeError(L.ref_last_error().decode())
    try:
        p = case["props"]
        f = p.get("f", (0.0, 0.0, 0.0))
        par = np.array([p["dt"], p["am"], p["af"], p["gam"], p["beta"], p["rho"], p.get("dmp", 0.0), f[0], f[1], f[2],
                        RefAssembly.ISO[p.get("iso", "nHook")], RefAssembly.VOL[p.get("vol")], p.get("C10", 0.0), p.get("C01", 0.0),
                        p.get("Kpen", 0.0), p.get("elM", 0.0), p.get("nu", 0.0)] + [0.0] * 8 + [100.0, 0.0, 0.0, 0.0], np.float64)
        Ag = _c(case["Ag"], np.float64); Yg = _c(case["Yg"], np.float64); Dg = _c(case["Dg"], np.float64)
        Bf = _c(case["Bf"], np.float64)
        Do = None if case.get("Do") is None else _c(case["Do"], np.float64)
        faces = case["faces"]
        f_info = np.array([[len(fa["nodes"]), fa["dof"], fa["bGrp"]] for fa in faces], np.int32).reshape(-1)
        f_nodes = np.concatenate([np.asarray(fa["nodes"], np.int32) for fa in faces])
        f_val = np.concatenate([np.asarray(fa["val"], np.float64).reshape(-1) for fa in faces])
        X = np.empty((m.nNo, 3)); out = np.zeros(9)
        ls = _c(ls, np.float64)
        incL = _c(case["incL"], np.int32); res = _c(case["res"], np.float64)
        kind = {"struct": 0, "lelas": 1, "mesh": 2}[case["kind"]]
        rc = L.ref_dropin_solid_step(h, int(mode), kind, Ag.shape[1], int(p.get("s", 0)), _p(par), _p(Ag), _p(Yg), _p(Dg), _p(Do),
                                     _p(Bf), len(faces), _p(f_info), _p(f_nodes), _p(f_val), _p(ls), _p(incL), _p(res), _p(X), _p(out))
        if rc != 0:
            raise RuntimeError(L.ref_last_error().decode())
    finally:
        L.ref_asm_destroy(h)
    keys = ("suc", "itr", "iNorm", "fNorm", "GM_itr", "CG_itr", "Resm", "Resc", "device_assembly")
    return X, dict(zip(keys, out))


def dropin_fluid_face_step(case, ls, mode, IENb, gE, hg, bfs=0.2):
    """dropin_fluid_step with one Neumann face assembled after the volume (set_bc_neu_l hook).  info["face_on_device"]."""
    L = dropin_lib()
    m = case["mesh"]
    x = _c(m.x, np.float64); ien = _c(m.ien, np.int32)
    h = L.ref_asm_create(m.nNo, m.nEl, ien.shape[1], _p(ien), _p(x), 1, -1.0)
    if not h:
        raise RuntimeError(L.ref_last_error().decode())
    try:
        p = case["props"]
        Ag = _c(case["Ag"], np.float64); Yg = _c(case["Yg"], np.float64); Bf = _c(case["Bf"], np.float64)
        visc = np.array([0, p["mu"], 0.0, 0.0, 0.0, 0.0], np.float64)
        faces = case["faces"]
        f_info = np.array([[len(f["nodes"]), f["dof"], f["bGrp"]] for f in faces], np.int32).reshape(-1)
        f_nodes = np.concatenate([np.asarray(f["nodes"], np.int32) for f in faces])
        f_val = np.concatenate([np.asarray(f["val"], np.float64).reshape(-1) for f in faces])
        X = np.empty((m.nNo, 4)); out = np.zeros(10)
        ls = _c(ls, np.float64)
        incL = _c(case["incL"], np.int32); res = _c(case["res"], np.float64)
        IENb = _c(IENb, np.int32); gE = _c(gE, np.int32); hg = _c(hg, np.float64)
        rc = L.ref_dropin_fluid_face_step(h, int(mode), Ag.shape[1], p["dt"], p["am"], p["af"], p["gam"], p["rho"], bfs, _p(visc),
                                          _p(Ag), _p(Yg), _p(Bf), len(faces), _p(f_info), _p(f_nodes), _p(f_val),
                                          IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), _p(hg), _p(ls), _p(incL), _p(res), _p(X), _p(out))
        if rc != 0:
            raise RuntimeError(L.ref_last_error().decode())
    finally:
        L.ref_asm_destroy(h)
    keys = ("suc", "itr", "iNorm", "fNorm", "GM_itr", "CG_itr", "Resm", "Resc", "device_assembly", "face_on_device")
    return X, dict(zip(keys, out))


def dropin_fluid_step(case, ls, mode):
    """One Newton-iteration hot path through the reference's own ComMod/eqType with
    eq.linear_algebra = B200LinearAlgebra.  mode 0: reference host assembly + device solve,
    mode 1: device assembly + device solve.  Returns (X (nNo,4), info dict)."""
    L = dropin_lib()
    m = case["mesh"]
    x = _c(m.x, np.float64); ien = _c(m.ien, np.int32)
    h = L.ref_asm_create(m.nNo, m.nEl, ien.shape[1], _p(ien), _p(x), 1, -1.0)
    if not h:
        raise RuntimeError(L.ref_last_error().decode())
    try:
        p = case["props"]
        Ag = _c(case["Ag"], np.float64); Yg = _c(case["Yg"], np.float64); Bf = _c(case["Bf"], np.float64)
        visc = np.array([p.get("viscType", 0), p["mu"], p.get("mu_o", 0.0), p.get("lam", 0.0), p.get("a", 0.0),
                         p.get("n", 0.0)], np.float64)
        fv = np.array(p.get("f", (0.0, 0.0, 0.0)), np.float64)
        faces = case["faces"]
        f_info = np.array([[len(f["nodes"]), f["dof"], f["bGrp"]] for f in faces], np.int32).reshape(-1)
        f_nodes = np.concatenate([np.asarray(f["nodes"], np.int32) for f in faces])
        f_val = np.concatenate([np.asarray(f["val"], np.float64).reshape(-1) for f in faces])
        X = np.empty((m.nNo, 4)); out = np.zeros(9)
        ls = _c(ls, np.float64)
        incL = _c(case["incL"], np.int32); res = _c(case["res"], np.float64)
        rc = L.ref_dropin_fluid_step(h, int(mode), Ag.shape[1], p["dt"], p["am"], p["af"], p["gam"], p["rho"], _p(fv),
                                     p.get("Kinv", 0.0), _p(visc), _p(Ag), _p(Yg), _p(Bf), len(faces), _p(f_info),
                                     _p(f_nodes), _p(f_val), _p(ls), _p(incL), _p(res), _p(X), _p(out))
        if rc != 0:
            raise RuntimeError(L.ref_last_error().decode())
    finally:
        L.ref_asm_destroy(h)
    keys = ("suc", "itr", "iNorm", "fNorm", "GM_itr", "CG_itr", "Resm", "Resc", "device_assembly")
    return X, dict(zip(keys, out))


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_asm_create.restype = C.c_void_p
        L.ref_asm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.ref_asm_destroy.argtypes = [C.c_void_p]
        L.ref_asm_nnz.argtypes = [C.c_void_p]
        L.ref_asm_get_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_asm_get_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_asm_fluid.restype = C.c_double
        L.ref_asm_fluid.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 5 + [C.c_void_p, C.c_double] + [C.c_void_p] * 6
        L.ref_asm_set_fibers.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_asm_solid.restype = C.c_double
        L.ref_asm_solid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8
        L.ref_asm_fsi.restype = C.c_double
        L.ref_asm_fsi.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 5 + [C.c_void_p] * 9
        L.ref_asm_ustruct.restype = C.c_double
        L.ref_asm_ustruct.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        L.ref_rank_create.restype = C.c_void_p
        L.ref_rank_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_rank_destroy.argtypes = [C.c_void_p]
        L.ref_rank_add_face.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_ranks_build.argtypes = [C.c_int, C.c_void_p]
        L.ref_rank_get_info.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ref_rank_get_req.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_ranks_solve.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_ranks_spmv.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_ranks_commuv.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_asm_domains.restype = C.c_double
        L.ref_asm_domains.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.ref_face_integ.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_fsi_ls_upd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]
        L.ref_pk2cc.argtypes = [C.c_void_p] * 6
        L.ref_pk2cc_dev.argtypes = [C.c_void_p] * 6
        L.ref_asm_bfolw.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8
        L.ref_asm_bneu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8
        L.ref_pic.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 12
        L.ref_asm_set_visc.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ref_asm_set_visc.restype = None
        L.ref_asm_set_prestress.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_asm_set_prestress.restype = None
        L.ref_asm_get_prestress.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_asm_get_prestress.restype = None
        L.ref_visc.argtypes = [C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 6
        L.ref_io_write_restart.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_io_history.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 7 + [C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class RefAssembly:
    """One mesh + one equation inside the reference's ComMod (see ref_harness.cpp)."""

    def __init__(self, x, ien, nFs: int = 1, qmTET4: float = -1.0):
        self.x = _c(x, np.float64)
        self.ien = _c(ien, np.int32)
        self.nNo = self.x.shape[0]
        self.nEl, self.eNoN = self.ien.shape
        self.h = lib().ref_asm_create(self.nNo, self.nEl, self.eNoN, _p(self.ien), _p(self.x), nFs, qmTET4)
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.nnz = lib().ref_asm_nnz(self.h)

    def close(self):
        if self.h:
            lib().ref_asm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def csr(self):
        rowPtr = np.empty(self.nNo + 1, np.int32)
        colPtr = np.empty(self.nnz, np.int32)
        lib().ref_asm_get_csr(self.h, _p(rowPtr), _p(colPtr))
        return rowPtr, colPtr

    def tables(self):
        nG = lib().ref_asm_get_tables(self.h, None, None, None)
        w = np.empty(nG)
        N = np.empty((nG, self.eNoN))
        Nx = np.empty((nG, self.eNoN, 3))
        lib().ref_asm_get_tables(self.h, _p(w), _p(N), _p(Nx))
        return w, N, Nx

    def fluid(self, Ag, Yg, Bf, *, dt, am, af, gam, rho, mu, f=(0.0, 0.0, 0.0), Kinv=0.0,
              visc=None, mvMsh=False):
        """construct_fluid (S/fluid.cpp:464).  Returns R (nNo,4), Val (nnz,16), seconds."""
        Ag = _c(Ag, np.float64); Yg = _c(Yg, np.float64); Bf = _c(Bf, np.float64)
        tDof = Ag.shape[1]
        v = np.array(visc if visc is not None else [0, mu, 0, 0, 0, 0], dtype=np.float64)
        fv = np.array(f, dtype=np.float64)
        R = np.empty((self.nNo, 4))
        Val = np.empty((self.nnz, 16))
        t = lib().ref_asm_fluid(self.h, tDof, int(mvMsh), dt, am, af, gam, rho, _p(fv), Kinv, _p(v),
                                _p(Ag), _p(Yg), _p(Bf), _p(R), _p(Val))
        if t < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return R, Val, t


    def struct_domains(self, elem_dmn, props, Ag, Yg, Dg, Bf):
        """construct_dsolid with eq.nDmn = len(props) struct domains; props: list of dicts as for solid()."""
        Ag = _c(Ag, np.float64); Yg = _c(Yg, np.float64); Dg = _c(Dg, np.float64); Bf = _c(Bf, np.float64)
        rows = []
        for p in props:
            f = p.get("f", (0.0, 0.0, 0.0)); ho = p.get("ho") or {}
            rows.append([p["dt"], p["am"], p["af"], p["gam"], p["beta"], p["rho"], p.get("dmp", 0.0), f[0], f[1], f[2],
                         self.ISO[p.get("iso", "nHook")], self.VOL[p.get("vol")], p.get("C10", 0.0), p.get("C01", 0.0), p.get("Kpen", 0.0),
                         0.0, 0.0] + [ho.get(k, 100.0 if k == "khs" else 0.0) for k in self.HO_KEYS] + [p.get("Tfa", 0.0), p.get("eta_s", 0.0), p.get("kap", 0.0)])
        par = np.array(rows, np.float64)
        ed = _c(elem_dmn, np.int32)
        # construct_dsolid reads the properties of com_mod.cDmn for every element (a copy where construct_fluid takes a
        # reference, see ref_harness.cpp): assemble one domain at a time and add, which is what the loop intends
        R = np.zeros((self.nNo, 3)); Val = np.zeros((self.nnz, 9))
        for d in range(len(props)):
            Rd = np.empty((self.nNo, 3)); Vd = np.empty((self.nnz, 9))
            t = lib().ref_asm_domains(self.h, 0, Ag.shape[1], len(props), par.shape[1], _p(par), _p(ed), d, _p(Ag), _p(Yg), _p(Dg), _p(Bf),
                                      _p(Rd), _p(Vd))
            if t < 0:
                raise RuntimeError(lib().ref_last_error().decode())
            R += Rd; Val += Vd
        return R, Val

    def fluid_domains(self, elem_dmn, props, Ag, Yg, Bf):
        """construct_fluid with eq.nDmn = len(props) fluid domains; props: list of dicts(dt, am, af, gam, rho, mu, f, Kinv, viscType...)."""
        Ag = _c(Ag, np.float64); Yg = _c(Yg, np.float64); Bf = _c(Bf, np.float64)
        rows = []
        for p in props:
            f = p.get("f", (0.0, 0.0, 0.0))
            rows.append([p["dt"], p["am"], p["af"], p["gam"], p["rho"], f[0], f[1], f[2], p.get("Kinv", 0.0), p.get("viscType", 0), p["mu"],
                         p.get("mu_o", 0.0), p.get("lam", 0.0), p.get("a", 0.0), p.get("n", 0.0), 0.0, 0.0])
        par = np.array(rows, np.float64)
        ed = _c(elem_dmn, np.int32)
        R = np.empty((self.nNo, 4)); Val = np.empty((self.nnz, 16))
        t = lib().ref_asm_domains(self.h, 3, Ag.shape[1], len(props), par.shape[1], _p(par), _p(ed), -1, _p(Ag), _p(Yg), None, _p(Bf), _p(R), _p(Val))
        if t < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return R, Val

    def face_integ(self, IENb, gE, s, l=0, u=None, geo=0, D=None):
        """all_fun::integ (S/all_fun.cpp:858) over one face: rows l..u of s (nNo, nrows); s None = the face area."""
        IENb = _c(IENb, np.int32); gE = _c(gE, np.int32)
        sa = None if s is None else _c(s, np.float64)
        Da = None if D is None else _c(D, np.float64)
        u = l if u is None else u
        out = C.c_double(0.0)
        rc = lib().ref_face_integ(self.h, IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), 0 if sa is None else sa.shape[1], _p(sa), l, u,
                                  geo, 0 if Da is None else Da.shape[1], _p(Da), C.byref(out))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return out.value

    def fsi_ls_upd(self, IENb, gE, gN, D, mvMsh=False):
        """eq_assem::fsi_ls_upd (S/eq_assem.cpp:316): val (nNoFace, 3) = int N_a n dGamma on x + Dn(0:2) or, moving mesh, x + Do(4:6)."""
        IENb = _c(IENb, np.int32); gE = _c(gE, np.int32); gN = _c(gN, np.int32); D = _c(D, np.float64)
        val = np.empty((len(gN), 3))
        rc = lib().ref_fsi_ls_upd(self.h, IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), len(gN), _p(gN), int(mvMsh), D.shape[1], _p(D), _p(val))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return val

    def bfolw(self, IENb, gE, hg, Dg, *, dt, af, beta=0.0, ustruct=False, am=1.0, gam=0.0):
        """eq_assem::b_neu_folw_p (S/eq_assem.cpp:186): follower pressure load on a struct face -> R (nNo,3), Val (nnz,9), or on a
        ustruct face -> R (nNo,4), Val (nnz,16), Kd (nnz,12)."""
        IENb = _c(IENb, np.int32); gE = _c(gE, np.int32); hg = _c(hg, np.float64); Dg = _c(Dg, np.float64)
        par = np.array([dt, af, beta, Dg.shape[1], float(ustruct), am, gam], np.float64)
        dof = 4 if ustruct else 3
        R = np.empty((self.nNo, dof)); Val = np.empty((self.nnz, dof * dof)); Kd = np.empty((self.nnz, 12))
        rc = lib().ref_asm_bfolw(self.h, IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), _p(par), _p(hg), _p(Dg), _p(R), _p(Val), _p(Kd))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return (R, Val, Kd) if ustruct else (R, Val)

    def bneu(self, kind, IENb, gE, hg, Yg, *, dt, af, gam, rho=0.0, bfs=0.0, mvMsh=False, Do=None):
        """b_assem_neu_bc (S/eq_assem.cpp:58) on one face: kind "fluid" (b_fluid, dof 4) or "solid" (b_l_elas, dof 3).
        Returns the face's contribution alone: R (nNo,dof), Val (nnz,dof*dof)."""
        IENb = _c(IENb, np.int32); gE = _c(gE, np.int32); hg = _c(hg, np.float64); Yg = _c(Yg, np.float64)
        Do = None if Do is None else _c(Do, np.float64)
        dof = 4 if kind == "fluid" else 3
        par = np.array([dt, af, gam, rho, bfs, Yg.shape[1], int(mvMsh)], np.float64)
        R = np.empty((self.nNo, dof)); Val = np.empty((self.nnz, dof * dof))
        rc = lib().ref_asm_bneu(self.h, 0 if kind == "fluid" else 1, IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE), _p(par), _p(hg),
                                _p(Yg), _p(Do), _p(R), _p(Val))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return R, Val

    def pk2cc(self, F, fl, *, iso="nHook", vol="ST91", C10=0.0, C01=0.0, Kpen=0.0, ho=None, Tfa=0.0, eta_s=0.0, kap=0.0, dev=False):
        """mat_models_carray::get_pk2cc<3> (S/mat_models_carray.h:182) or, dev=True, mat_models::get_pk2cc_dev (S/mat_models.cpp:630,
        the ustruct form without volumetric terms): S (3,3) and Dm (6,6) at the deformation gradient F."""
        ho = ho or {}
        par = np.array([0.0] * 10 + [self.ISO[iso], self.VOL[vol], C10, C01, Kpen, 0.0, 0.0]
                       + [ho.get(k, 100.0 if k == "khs" else 0.0) for k in self.HO_KEYS] + [Tfa, eta_s, kap], np.float64)
        F = _c(F, np.float64); fl = _c(fl, np.float64)
        S = np.empty((3, 3)); Dm = np.empty((6, 6))
        rc = (lib().ref_pk2cc_dev if dev else lib().ref_pk2cc)(self.h, _p(par), _p(F), _p(fl), _p(S), _p(Dm))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return S, Dm

    ISO = {"nHook": 0, "StVK": 1, "mStVK": 2, "HO": 3, "MR": 4, "HGO": 5, "Gucci": 6, "HO_ma": 7}
    HO_KEYS = ("a", "b", "aff", "bff", "ass", "bss", "afs", "bfs", "khs")

    def set_fibers(self, fN):
        """lM.fN: (nEl, 6) fibre + sheet directions, or None."""
        if fN is None:
            lib().ref_asm_set_fibers(self.h, 0, None)
        else:
            fN = _c(fN, np.float64)
            lib().ref_asm_set_fibers(self.h, 2, _p(fN))
    VOL = {None: 0, "Quad": 1, "ST91": 2, "M94": 3}

    def solid(self, kind, Ag, Yg, Dg, Bf, *, dt, am, af, gam, beta, rho, dmp=0.0, f=(0.0, 0.0, 0.0), iso="nHook",
              vol="ST91", C10=0.0, C01=0.0, Kpen=0.0, elM=0.0, nu=0.0, s=0, Do=None, ho=None, Tfa=0.0, eta_s=0.0, kap=0.0,
              visc=None, visc_mu=0.0, pS0=None, pstEq=False):
        """kind "struct": construct_dsolid (S/sv_struct.cpp:213); "lelas": construct_l_elas (S/l_elas.cpp:58).
        Returns R (nNo,3), Val (nnz,9), seconds."""
        Ag = _c(Ag, np.float64); Yg = _c(Yg, np.float64); Dg = _c(Dg, np.float64); Bf = _c(Bf, np.float64)
        tDof = Ag.shape[1]
        ho = ho or {}
        par = np.array([dt, am, af, gam, beta, rho, dmp, f[0], f[1], f[2], self.ISO[iso], self.VOL[vol], C10, C01, Kpen,
                        elM, nu] + [ho.get(k, 100.0 if k == "khs" else 0.0) for k in self.HO_KEYS] + [Tfa, eta_s, kap], np.float64)
        R = np.empty((self.nNo, 3))
        Val = np.empty((self.nnz, 9))
        Do = None if Do is None else _c(Do, np.float64)
        lib().ref_asm_set_visc(self.h, {None: 0, "newt": 1, "pot": 2}[visc], float(visc_mu))      # dmn.solid_visc
        pS0 = None if pS0 is None else _c(pS0, np.float64)
        lib().ref_asm_set_prestress(self.h, None if pS0 is None else _p(pS0), int(pstEq))             # com_mod.pS0 / pstEq
        t = lib().ref_asm_solid(self.h, {"struct": 0, "lelas": 1, "mesh": 2}[kind], tDof, int(s), _p(par), _p(Ag), _p(Yg),
                                _p(Dg), _p(Do), _p(Bf), _p(R), _p(Val))
        if t < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        self.pSn = self.pSa = None
        if pstEq:
            self.pSn = np.empty((self.nNo, 6)); self.pSa = np.empty(self.nNo)
            lib().ref_asm_get_prestress(self.h, _p(self.pSn), _p(self.pSa))
        return R, Val, t


    def fsi(self, elem_dmn, Ag, Yg, Dg, Bf, *, dt, am, af, gam, beta, fluid, solid, pS0=None):
        """construct_fsi (S/fsi.cpp:42) with domain 0 = fluid, 1 = struct.  fluid: dict(rho, mu, f, viscType...);
        solid: dict(rho, dmp, f, iso, vol, C10, C01, Kpen).  Returns R (nNo,4), Val (nnz,16), seconds."""
        Ag = _c(Ag, np.float64); Yg = _c(Yg, np.float64); Dg = _c(Dg, np.float64); Bf = _c(Bf, np.float64)
        ed = _c(elem_dmn, np.int32)
        ff = fluid.get("f", (0.0, 0.0, 0.0)); sf = solid.get("f", (0.0, 0.0, 0.0))
        fpar = np.array([fluid["rho"], ff[0], ff[1], ff[2], fluid.get("viscType", 0), fluid["mu"], fluid.get("mu_o", 0.0),
                         fluid.get("lam", 0.0), fluid.get("a", 0.0), fluid.get("n", 0.0)], np.float64)
        spar = np.array([dt, am, af, gam, beta, solid["rho"], solid.get("dmp", 0.0), sf[0], sf[1], sf[2],
                         self.ISO[solid.get("iso", "nHook")], self.VOL[solid.get("vol", "ST91")], solid["C10"],
                         solid.get("C01", 0.0), solid.get("Kpen", 0.0), 0.0, 0.0] + [0.0] * 8 + [100.0, 0.0, 0.0, 0.0], np.float64)
        R = np.empty((self.nNo, 4))
        Val = np.empty((self.nnz, 16))
        # wall viscosity (solid["visc"], solid["visc_mu"]) and prestress (com_mod.pS0; construct_fsi never accumulates pSn / pSa)
        lib().ref_asm_set_visc(self.h, {None: 0, "newt": 1, "pot": 2}[solid.get("visc")], float(solid.get("visc_mu", 0.0)))
        pS0 = None if pS0 is None else _c(pS0, np.float64)
        lib().ref_asm_set_prestress(self.h, None if pS0 is None else _p(pS0), 0)
        t = lib().ref_asm_fsi(self.h, Ag.shape[1], dt, am, af, gam, beta, _p(fpar), _p(spar), _p(ed), _p(Ag), _p(Yg), _p(Dg),
                              _p(Bf), _p(R), _p(Val))
        if t < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return R, Val, t


    def ustruct(self, Ag, Yg, Dg, Bf, *, dt, am, af, gam, rho, elM, nu, ctM, ctC, vol, C10, Kpen, f=(0.0, 0.0, 0.0), Ad=None,
                iso="nHook", ho=None, Tfa=0.0, eta_s=0.0, C01=0.0, kap=0.0, visc=None, visc_mu=0.0, **_ignored):
        """construct_usolid (S/ustruct.cpp:216) [+ ustruct_r when Ad is given].  Returns R (nNo,4), Val (nnz,16),
        Kd (nnz,12), seconds."""
        Ag = _c(Ag, np.float64); Yg = _c(Yg, np.float64); Dg = _c(Dg, np.float64); Bf = _c(Bf, np.float64)
        Ad = None if Ad is None else _c(Ad, np.float64)
        ho = ho or {}
        par = np.array([dt, am, af, gam, rho, f[0], f[1], f[2], elM, nu, ctM, ctC, self.VOL[vol], C10, Kpen, self.ISO[iso]]
                       + [ho.get(k, 100.0 if k == "khs" else 0.0) for k in self.HO_KEYS] + [Tfa, eta_s, C01, kap], np.float64)
        R = np.empty((self.nNo, 4)); Val = np.empty((self.nnz, 16)); Kd = np.empty((self.nnz, 12))
        lib().ref_asm_set_visc(self.h, {None: 0, "newt": 1, "pot": 2}[visc], float(visc_mu))      # dmn.solid_visc
        t = lib().ref_asm_ustruct(self.h, Ag.shape[1], _p(par), _p(Ag), _p(Yg), _p(Dg), _p(Bf), _p(Ad), _p(R), _p(Val), _p(Kd))
        if t < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return R, Val, Kd, t


def ls_params(ls_type, relTol, absTol=1e-10, mItr=10, sD=100, gm=(1e-2, 1e-10, 2, 100), cg=(0.2, 1e-10, 500)):
    return np.array([ls_type, relTol, absTol, mItr, sD, gm[0], gm[1], gm[2], gm[3], cg[0], cg[1], cg[2]], np.float64)


OUT_FIELDS = ("suc", "itr", "iNorm", "fNorm", "dB", "callD", "GM_itr", "CG_itr", "Resm", "Resc",
              "GM_callD", "CG_callD", "wall_s")


class RefRanks:
    """nranks "MPI ranks" (threads) of the reference FSILS: lhs_create + bc_create + solve."""

    def __init__(self, parts):
        """parts: list of dict(gnNo, gNodes, rowPtr, colPtr, faces=[dict(nodes(local ids), dof, bGrp, val|None)])"""
        L = lib()
        self.n = len(parts)
        self.parts = parts
        self.hs = []
        self._keep = []
        for p in parts:
            g = _c(p["gNodes"], np.int32); r = _c(p["rowPtr"], np.int32); c = _c(p["colPtr"], np.int32)
            h = L.ref_rank_create(int(p["gnNo"]), len(g), len(c), _p(g), _p(r), _p(c))
            for f in p.get("faces", []):
                nodes = _c(f["nodes"], np.int32)
                val = None if f.get("val") is None else _c(f["val"], np.float64)
                L.ref_rank_add_face(h, len(nodes), int(f["dof"]), int(f["bGrp"]), _p(nodes), _p(val))
            self.hs.append(h)
        self.harr = (C.c_void_p * self.n)(*self.hs)
        if L.ref_ranks_build(self.n, self.harr) != 0:
            raise RuntimeError(L.ref_last_error().decode())

    def close(self):
        for h in self.hs:
            lib().ref_rank_destroy(h)
        self.hs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, p):
        nNo = len(self.parts[p]["gNodes"]); nnz = len(self.parts[p]["colPtr"])
        info = np.zeros(3, np.int32); mp = np.empty(nNo, np.int32); rp = np.empty((nNo, 2), np.int32)
        cp = np.empty(nnz, np.int32); dp = np.empty(nNo, np.int32)
        lib().ref_rank_get_info(self.hs[p], _p(info), _p(mp), _p(rp), _p(cp), _p(dp))
        reqs = []
        for i in range(int(info[2])):
            iP = C.c_int(0)
            n = lib().ref_rank_get_req(self.hs[p], i, C.byref(iP), None)
            ptr = np.empty(n, np.int32)
            lib().ref_rank_get_req(self.hs[p], i, C.byref(iP), _p(ptr))
            reqs.append((iP.value, ptr))
        return dict(mynNo=int(info[0]), shnNo=int(info[1]), nReq=int(info[2]), map=mp, rowPtr=rp, colPtr=cp,
                    diagPtr=dp, reqs=reqs)

    def _ptrs(self, arrs):
        return (C.c_void_p * self.n)(*[a.ctypes.data for a in arrs])

    def solve(self, dof, ls, prec, R, Val, incL=None, res=None):
        """fsils_solve (L/solve.cpp:50) on every rank.  R, Val: lists of per-rank arrays (copied).
        Returns (X list, Val_scaled list, out list[dict])."""
        Rs = [_c(r, np.float64).copy() for r in R]
        Vs = [_c(v, np.float64).copy() for v in Val]
        out = np.zeros((self.n, len(OUT_FIELDS)))
        incL_a = None if incL is None else _c(incL, np.int32)
        res_a = None if res is None else _c(res, np.float64)
        ls = _c(ls, np.float64)
        rc = lib().ref_ranks_solve(self.n, self.harr, dof, _p(ls), int(prec), self._ptrs(Rs), self._ptrs(Vs),
                                   _p(incL_a), _p(res_a), _p(out))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return Rs, Vs, [dict(zip(OUT_FIELDS, o)) for o in out]

    def spmv(self, dof, Val, X, reps=1):
        Vs = [_c(v, np.float64) for v in Val]
        Xs = [_c(x, np.float64) for x in X]
        Ys = [np.zeros_like(x) for x in Xs]
        secs = C.c_double(0.0)
        rc = lib().ref_ranks_spmv(self.n, self.harr, dof, self._ptrs(Vs), self._ptrs(Xs), self._ptrs(Ys), reps,
                                  C.byref(secs))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return Ys, secs.value

    def commuv(self, dof, V):
        Vs = [_c(v, np.float64).copy() for v in V]
        lib().ref_ranks_commuv(self.n, self.harr, dof, self._ptrs(Vs))
        return Vs


PHYS = dict(fluid=0, struct=1, lElas=2, ustruct=3, FSI=4, mesh=5)


def pic(op, state, eqs, *, dt, cEq=0, dFlag=False, sstEq=False, R=None, Rd=None):
    """The reference's pic::picp / pici / picc (S/pic.cpp:591,486,74) on host arrays.  op: "p", "i" or "c".
    state: dict with Ao, Yo, Do, An, Yn, Dn (nNo,tDof), Ad (nNo,3), Ag, Yg, Dg; modified copies are returned.
    eqs: list of dict(s, e, am, af, gam, beta, phys, itr)."""
    st = {k: _c(v, np.float64).copy() for k, v in state.items()}
    nNo, tDof = st["Ao"].shape
    par = np.array([[q["s"], q["e"], q["am"], q["af"], q["gam"], q.get("beta", 0.0), PHYS[q["phys"]], q.get("itr", 1)] for q in eqs],
                   np.float64)
    Ra = None if R is None else _c(R, np.float64)
    Rda = None if Rd is None else _c(Rd, np.float64)
    rc = lib().ref_pic({"p": 0, "i": 1, "c": 2}[op], nNo, tDof, len(eqs), cEq, _p(par), int(dFlag), int(sstEq), dt,
                       *[_p(st[k]) for k in ("Ao", "Yo", "Do", "An", "Yn", "Dn", "Ad", "Ag", "Yg", "Dg")], _p(Ra), _p(Rda))
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return st


def io_write_restart(stem, *, stamp, cTS, time, iNorm, xn, Yn, An, Dn=None, Ad=None, pS0=None):
    """output::write_restart (S/output.cpp:202-345) for one rank; arrays (tnNo, tDof) row-major = the solver's column-major buffers.
    Returns the record length the reference computes (S/initialize.cpp:505-513).  The file is <stem>_NNN.bin."""
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
    st = np.ascontiguousarray(stamp, np.int32)
    iNorm, xn, Yn, An, Dn, Ad, pS0 = map(f64, (iNorm, xn, Yn, An, Dn, Ad, pS0))
    recLn = C.c_longlong()
    rc = lib().ref_io_write_restart(os.fsencode(stem), _p(st), int(cTS), float(time), iNorm.size, _p(iNorm), xn.size, _p(xn),
                                    Yn.shape[1], Yn.shape[0], _p(Yn), _p(An), None if Dn is None else _p(Dn),
                                    0 if Ad is None else Ad.shape[1], None if Ad is None else _p(Ad),
                                    0 if pS0 is None else pS0.shape[1], None if pS0 is None else _p(pS0), C.byref(recLn))
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return recLn.value


def io_history(path, *, nEq, sym, cTS, itr, saved, elapsed, eq_iNorm, eq_pNorm, ri_iNorm, ri_fNorm, ri_dB, ri_callD, ri_itr, ri_suc):
    """output::output_result (S/output.cpp:46-180): header block + one line into `path`; returns the file's text."""
    rc = lib().ref_io_history(os.fsencode(path), nEq, sym.encode(), cTS, itr, int(saved), elapsed, eq_iNorm, eq_pNorm, ri_iNorm, ri_fNorm,
                              ri_dB, ri_callD, ri_itr, int(ri_suc))
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())
    with open(path) as f:
        return f.read()


def visc(model, mu, Nx, vx, F):
    """mat_models_carray::get_visc_stress_and_tangent<3> (S/mat_models_carray.h:1578): model "newt" | "pot"; Nx (eNoN, 3).
    Returns Svis (3,3), Kvis_u, Kvis_v as (eNoN_a, eNoN_b, 3, 3)."""
    Nx = np.ascontiguousarray(Nx, np.float64); vx = np.ascontiguousarray(vx, np.float64); F = np.ascontiguousarray(F, np.float64)
    n = Nx.shape[0]
    S = np.zeros((3, 3)); Ku = np.zeros((n, n, 9)); Kv = np.zeros((n, n, 9))       # Array3(9, a, b): memory order [b][a][ii]
    rc = lib().ref_visc({"newt": 1, "pot": 2}[model], float(mu), n, _p(Nx), _p(vx), _p(F), _p(S), _p(Ku), _p(Kv))
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return S, Ku.transpose(1, 0, 2).reshape(n, n, 3, 3), Kv.transpose(1, 0, 2).reshape(n, n, 3, 3)
