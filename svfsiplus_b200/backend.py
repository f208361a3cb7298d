"""ctypes binding of the C ABI in include/svb200.h (svfsiplus_b200/libsvb200.so).

This is the Python-side stub of the drop-in boundary, used by the parity tests, bench.py and
__graft_entry__.smoke().  It mirrors, call for call, what the C++ plug-in class
svfsiplus_b200/host/B200LinearAlgebra.cpp does inside svMultiPhysics (LinearAlgebra interface,
Code/Source/solver/LinearAlgebra.h:39-63).  There is no fallback: if the CUDA library is missing or
no device is present every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvb200.so")

LS_BICGS, LS_NS, LS_GMRES, LS_CG = 795, 796, 797, 798
PREC_FSILS, PREC_RCS = 701, 709
BC_DIR, BC_NEU = 0, 1


class Tol(C.Structure):
    _fields_ = [("relTol", C.c_double), ("absTol", C.c_double), ("mItr", C.c_int), ("sD", C.c_int)]


class SubOut(C.Structure):
    _fields_ = [("suc", C.c_int), ("itr", C.c_int), ("iNorm", C.c_double), ("fNorm", C.c_double),
                ("dB", C.c_double), ("callD", C.c_double)]


class LsOut(C.Structure):
    _fields_ = [("RI", SubOut), ("GM", SubOut), ("CG", SubOut), ("Resm", C.c_int), ("Resc", C.c_int)]


class FluidProps(C.Structure):
    _fields_ = [("dt", C.c_double), ("am", C.c_double), ("af", C.c_double), ("gam", C.c_double),
                ("tDof", C.c_int), ("mvMsh", C.c_int),
                ("rho", C.c_double), ("f", C.c_double * 3), ("Kinv", C.c_double),
                ("viscType", C.c_int),
                ("mu_i", C.c_double), ("mu_o", C.c_double), ("lam", C.c_double), ("a", C.c_double), ("n", C.c_double)]


class StructProps(C.Structure):
    _fields_ = [("dt", C.c_double), ("am", C.c_double), ("af", C.c_double), ("gam", C.c_double), ("beta", C.c_double),
                ("tDof", C.c_int), ("s", C.c_int),
                ("rho", C.c_double), ("dmp", C.c_double), ("f", C.c_double * 3),
                ("isoType", C.c_int), ("volType", C.c_int),
                ("C10", C.c_double), ("C01", C.c_double), ("Kpen", C.c_double),
                ("a", C.c_double), ("b", C.c_double), ("aff", C.c_double), ("bff", C.c_double), ("ass", C.c_double),
                ("bss", C.c_double), ("afs", C.c_double), ("bfs", C.c_double), ("khs", C.c_double),
                ("Tfa", C.c_double), ("Tsa", C.c_double), ("kap", C.c_double),
                ("viscType", C.c_int), ("visc_mu", C.c_double)]


class LelasProps(C.Structure):
    _fields_ = [("dt", C.c_double), ("am", C.c_double), ("af", C.c_double), ("beta", C.c_double),
                ("tDof", C.c_int), ("s", C.c_int), ("mesh_mode", C.c_int),
                ("rho", C.c_double), ("elM", C.c_double), ("nu", C.c_double), ("f", C.c_double * 3)]


class UstructProps(C.Structure):
    _fields_ = [("dt", C.c_double), ("am", C.c_double), ("af", C.c_double), ("gam", C.c_double),
                ("tDof", C.c_int), ("s", C.c_int),
                ("rho", C.c_double), ("f", C.c_double * 3),
                ("elM", C.c_double), ("nu", C.c_double), ("ctM", C.c_double), ("ctC", C.c_double),
                ("isoType", C.c_int), ("volType", C.c_int),
                ("C10", C.c_double), ("Kpen", C.c_double),
                ("a", C.c_double), ("b", C.c_double), ("aff", C.c_double), ("bff", C.c_double), ("ass", C.c_double),
                ("bss", C.c_double), ("afs", C.c_double), ("bfs", C.c_double), ("khs", C.c_double),
                ("Tfa", C.c_double), ("Tsa", C.c_double), ("C01", C.c_double), ("kap", C.c_double),
                ("viscType", C.c_int), ("visc_mu", C.c_double)]


class BneuProps(C.Structure):
    _fields_ = [("dt", C.c_double), ("af", C.c_double), ("gam", C.c_double), ("tDof", C.c_int), ("mvMsh", C.c_int),
                ("rho", C.c_double), ("bfs", C.c_double)]


class BfolwProps(C.Structure):
    _fields_ = [("dt", C.c_double), ("af", C.c_double), ("beta", C.c_double), ("tDof", C.c_int), ("s", C.c_int),
                ("ustruct", C.c_int), ("am", C.c_double), ("gam", C.c_double)]


class PicEq(C.Structure):
    _fields_ = [("s", C.c_int), ("e", C.c_int), ("am", C.c_double), ("af", C.c_double), ("gam", C.c_double),
                ("beta", C.c_double), ("kind", C.c_int)]


PIC = dict(Ao=0, Yo=1, Do=2, An=3, Yn=4, Dn=5, Ad=6, Ag=7, Yg=8, Dg=9)

ISO_TYPES = {"nHook": 0, "StVK": 1, "mStVK": 2, "HO": 3, "MR": 4, "HGO": 5, "Gucci": 6, "HO_ma": 7}
VOL_TYPES = {None: 0, "Quad": 1, "ST91": 2, "M94": 3}

EXPORTS = [
    "b200_create", "b200_destroy", "b200_last_error", "b200_device_count", "b200_elem_tables", "b200_comm_unique_id", "b200_comm_init", "b200_comm_transport",
    "b200_lhs_create", "b200_face_set", "b200_mesh_set", "b200_zero", "b200_state_set", "b200_assemble_fluid",
    "b200_disp_set", "b200_assemble_struct", "b200_assemble_lelas", "b200_mesh_domains", "b200_mesh_fibers", "b200_assemble_fsi",
    "b200_assemble_ustruct", "b200_ustruct_r", "b200_get_Kd",
    "b200_assemble_elem", "b200_get_R", "b200_set_R", "b200_add_R", "b200_get_Val", "b200_set_Val", "b200_commu_R", "b200_solve",
    "b200_spmv", "b200_op_bench", "b200_launch_count", "b200_tune", "b200_last_timings", "b200_profile", "b200_profile_read",
    "b200_timer",
    "b200_pic_init", "b200_pic_set", "b200_pic_get", "b200_pic_scatter", "b200_picp", "b200_pici", "b200_picc",
    "b200_pic_copy_rows", "b200_pic_advance", "b200_face_mesh_set", "b200_assemble_bneu",
    "b200_assemble_fluid_dmn", "b200_assemble_struct_dmn",
    "b200_assemble_bfolw", "b200_face_integ", "b200_face_normal_update", "b200_face_get_val", "b200_pattern_begin", "b200_pattern_add_mesh", "b200_pattern_finish", "b200_pattern_get",
    "b200_lhs_layout_create", "b200_lhs_layout_sizes", "b200_lhs_layout_map", "b200_lhs_layout_req", "b200_lhs_layout_free",
    "b200_partition_rcb", "b200_partition_metis", "b200_prestress_set", "b200_prestress_get",
]

KERNEL_CLASSES = ["spmv_vv4", "spmv_vv3", "spmv_ss", "spmv_sv", "spmv_vs", "multi_dot", "cgs_update_scale", "blas1",
                  "scale_val", "depart", "assembly", "halo"]

_lib = None


def lib():
    """Load libsvb200.so (fails loudly when the CUDA extension has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.b200_create.argtypes = [C.POINTER(vp), ci]
        L.b200_destroy.argtypes = [vp]
        L.b200_destroy.restype = None
        L.b200_last_error.argtypes = [vp]
        L.b200_last_error.restype = C.c_char_p
        L.b200_elem_tables.argtypes = [ci, cd, vp, vp, vp]
        L.b200_comm_unique_id.argtypes = [vp]
        L.b200_comm_init.argtypes = [vp, ci, ci, vp]
        L.b200_comm_transport.argtypes = [vp]
        L.b200_comm_transport.restype = C.c_char_p
        L.b200_lhs_create.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, ci, vp, vp, vp, ci]
        L.b200_lhs_layout_create.argtypes = [ci, ci, ci, vp, vp, C.POINTER(vp)]
        L.b200_lhs_layout_sizes.argtypes = [vp, vp, vp, vp, vp]
        L.b200_lhs_layout_map.argtypes = [vp, vp]
        L.b200_lhs_layout_req.argtypes = [vp, ci, vp, vp, vp]
        L.b200_lhs_layout_free.argtypes = [vp]
        L.b200_lhs_layout_free.restype = None
        L.b200_partition_rcb.argtypes = [ci, vp, ci, vp]
        L.b200_prestress_set.argtypes = [vp, vp, ci]
        L.b200_prestress_get.argtypes = [vp, vp, vp]
        L.b200_face_set.argtypes = [vp, ci, ci, ci, ci, vp, vp, ci]
        L.b200_mesh_set.argtypes = [vp, ci, ci, vp, vp, cd]
        L.b200_zero.argtypes = [vp, ci]
        L.b200_state_set.argtypes = [vp, ci, vp, vp, vp]
        L.b200_assemble_fluid.argtypes = [vp, C.POINTER(FluidProps)]
        L.b200_disp_set.argtypes = [vp, ci, vp, vp]
        L.b200_assemble_struct.argtypes = [vp, C.POINTER(StructProps)]
        L.b200_assemble_lelas.argtypes = [vp, C.POINTER(LelasProps)]
        L.b200_assemble_ustruct.argtypes = [vp, C.POINTER(UstructProps)]
        L.b200_ustruct_r.argtypes = [vp, cd, cd, ci, vp]
        L.b200_get_Kd.argtypes = [vp, vp]
        L.b200_mesh_domains.argtypes = [vp, ci, vp]
        L.b200_mesh_fibers.argtypes = [vp, ci, vp]
        L.b200_assemble_fsi.argtypes = [vp, ci, vp, C.POINTER(FluidProps), C.POINTER(StructProps)]
        L.b200_assemble_elem.argtypes = [vp, ci, vp, vp, vp]
        L.b200_get_R.argtypes = [vp, vp]
        L.b200_set_R.argtypes = [vp, ci, vp]
        L.b200_add_R.argtypes = [vp, ci, vp]
        L.b200_get_Val.argtypes = [vp, vp]
        L.b200_set_Val.argtypes = [vp, ci, vp]
        L.b200_commu_R.argtypes = [vp]
        L.b200_solve.argtypes = [vp, ci, ci, C.POINTER(Tol), C.POINTER(Tol), C.POINTER(Tol), vp, vp, vp, C.POINTER(LsOut)]
        L.b200_spmv.argtypes = [vp, ci, vp, vp]
        L.b200_op_bench.argtypes = [vp, ci, ci, ci, C.POINTER(cd), C.POINTER(cd)]
        L.b200_launch_count.argtypes = [vp]
        L.b200_launch_count.restype = C.c_longlong
        L.b200_last_timings.argtypes = [vp, vp]
        L.b200_profile.argtypes = [vp, ci]
        L.b200_profile_read.argtypes = [vp, ci, vp, vp, vp]
        L.b200_timer.argtypes = [vp, ci, C.POINTER(cd)]
        L.b200_pic_init.argtypes = [vp, ci, ci, C.POINTER(PicEq), ci, ci]
        L.b200_pic_set.argtypes = [vp, ci, vp]
        L.b200_pic_get.argtypes = [vp, ci, vp]
        L.b200_pic_scatter.argtypes = [vp, ci, ci, vp, vp]
        L.b200_picp.argtypes = [vp, cd]
        L.b200_pici.argtypes = [vp]
        L.b200_picc.argtypes = [vp, ci, cd, ci]
        L.b200_pic_copy_rows.argtypes = [vp, ci, vp, ci, ci]
        L.b200_pic_advance.argtypes = [vp]
        L.b200_pattern_begin.argtypes = [vp, ci]
        L.b200_pattern_add_mesh.argtypes = [vp, ci, ci, vp]
        L.b200_pattern_finish.argtypes = [vp, C.POINTER(ci)]
        L.b200_pattern_get.argtypes = [vp, vp, vp]
        L.b200_assemble_fluid_dmn.argtypes = [vp, ci, C.POINTER(FluidProps)]
        L.b200_assemble_struct_dmn.argtypes = [vp, ci, C.POINTER(StructProps)]
        L.b200_face_mesh_set.argtypes = [vp, ci, ci, ci, vp, vp]
        L.b200_face_integ.argtypes = [vp, ci, ci, ci, ci, ci, C.POINTER(cd)]
        L.b200_face_normal_update.argtypes = [vp, ci, ci, ci]
        L.b200_face_get_val.argtypes = [vp, ci, vp]
        L.b200_assemble_bfolw.argtypes = [vp, ci, C.POINTER(BfolwProps), vp]
        L.b200_assemble_bneu.argtypes = [vp, ci, ci, C.POINTER(BneuProps), vp]
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))          # raw address (e.g. a pinned torch tensor's data_ptr())


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def fluid_props(*, dt, am, af, gam, rho, mu, tDof=4, mvMsh=False, f=(0.0, 0.0, 0.0), Kinv=0.0,
                viscType=0, mu_o=0.0, lam=0.0, a=0.0, n=0.0) -> FluidProps:
    p = FluidProps()
    p.dt, p.am, p.af, p.gam = dt, am, af, gam
    p.tDof, p.mvMsh = tDof, int(mvMsh)
    p.rho, p.Kinv = rho, Kinv
    p.f[0], p.f[1], p.f[2] = f
    p.viscType, p.mu_i, p.mu_o, p.lam, p.a, p.n = viscType, mu, mu_o, lam, a, n
    return p


def struct_props(*, dt, am, af, gam, beta, rho, tDof=3, s=0, dmp=0.0, f=(0.0, 0.0, 0.0), iso="nHook", vol="ST91",
                 C10=0.0, C01=0.0, Kpen=0.0, ho=None, Tfa=0.0, eta_s=0.0, kap=0.0, visc=None, visc_mu=0.0) -> StructProps:
    p = StructProps()
    p.dt, p.am, p.af, p.gam, p.beta = dt, am, af, gam, beta
    p.tDof, p.s = tDof, s
    p.rho, p.dmp = rho, dmp
    p.f[0], p.f[1], p.f[2] = f
    p.isoType, p.volType = ISO_TYPES[iso], VOL_TYPES[vol]
    p.C10, p.C01, p.Kpen = C10, C01, Kpen
    p.Tfa, p.Tsa, p.kap = Tfa, Tfa * eta_s, kap
    p.viscType, p.visc_mu = {None: 0, "newt": 1, "pot": 2}[visc], visc_mu
    p.khs = 100.0
    for k, v in (ho or {}).items():
        setattr(p, k, v)
    return p


def lelas_props(*, dt, am, af, beta, rho, elM, nu, tDof=3, s=0, f=(0.0, 0.0, 0.0), mesh_mode=False, **_ignored) -> LelasProps:
    p = LelasProps()
    p.dt, p.am, p.af, p.beta = dt, am, af, beta
    p.tDof, p.s, p.mesh_mode = tDof, s, int(mesh_mode)
    p.rho, p.elM, p.nu = rho, elM, nu
    p.f[0], p.f[1], p.f[2] = f
    return p


def ustruct_props(*, dt, am, af, gam, rho, elM, nu, ctM, ctC, vol, C10, Kpen, tDof=4, s=0, f=(0.0, 0.0, 0.0), iso="nHook",
                  ho=None, Tfa=0.0, eta_s=0.0, C01=0.0, kap=0.0, visc=None, visc_mu=0.0, **_ignored) -> UstructProps:
    p = UstructProps()
    p.dt, p.am, p.af, p.gam = dt, am, af, gam
    p.tDof, p.s = tDof, s
    p.rho = rho
    p.f[0], p.f[1], p.f[2] = f
    p.elM, p.nu, p.ctM, p.ctC = elM, nu, ctM, ctC
    p.isoType, p.volType = ISO_TYPES[iso], VOL_TYPES[vol]
    p.C10, p.Kpen = C10, Kpen
    p.Tfa, p.Tsa, p.C01, p.kap = Tfa, Tfa * eta_s, C01, kap
    p.viscType, p.visc_mu = {None: 0, "newt": 1, "pot": 2}[visc], visc_mu
    p.khs = 100.0
    for k, v in (ho or {}).items():
        setattr(p, k, v)
    return p


def elem_tables(eNoN, qmTET4=-1.0):
    """(w[nG], N[nG,eNoN], Nxi[nG,eNoN,3]) of the element kernels; host-only."""
    nG = {4: 4, 8: 8}[eNoN]
    w = np.zeros(nG); N = np.zeros((nG, eNoN)); Nxi = np.zeros((nG, eNoN, 3))
    if lib().b200_elem_tables(eNoN, qmTET4, _p(w), _p(N), _p(Nxi)) != nG:
        raise RuntimeError("b200_elem_tables failed")
    return w, N, Nxi


def sub_out_dict(s: SubOut):
    return dict(suc=bool(s.suc), itr=s.itr, iNorm=s.iNorm, fNorm=s.fNorm, dB=s.dB, callD=s.callD)


def unique_id() -> np.ndarray:
    """128-byte NCCL unique id, made on rank 0 and distributed by the caller (torch.distributed / MPI_Bcast)."""
    uid = np.zeros(128, np.uint8)
    if lib().b200_comm_unique_id(_p(uid)) != 0:
        raise RuntimeError("b200_comm_unique_id: " + lib().b200_last_error(None).decode())
    return uid


def partition_rcb(centroids, nparts: int) -> np.ndarray:
    """Element -> rank map by recursive coordinate bisection (b200_partition_rcb; host side, no device)."""
    c = _c(centroids, np.float64)
    part = np.zeros(c.shape[0], np.int32)
    if lib().b200_partition_rcb(c.shape[0], _p(c), int(nparts), _p(part)) != 0:
        raise RuntimeError("b200_partition_rcb: " + lib().b200_last_error(None).decode())
    return part


def partition_metis(ien, nNo: int, nparts: int, ncommon: int):
    """Element -> rank map by METIS' k-way partition of the dual graph (b200_partition_metis; the reference's criterion,
    distribute.cpp:1683-1706: ncommon = nodes of a boundary element).  Returns (part, edgecut)."""
    ien = _c(ien, np.int32)
    part = np.zeros(ien.shape[0], np.int32)
    cut = C.c_longlong(0)
    L = lib()
    L.b200_partition_metis.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_longlong)]
    if L.b200_partition_metis(ien.shape[0], ien.shape[1], int(nNo), _p(ien), int(ncommon), int(nparts), _p(part), C.byref(cut)) != 0:
        raise RuntimeError("b200_partition_metis: " + L.b200_last_error(None).decode())
    return part, int(cut.value)


def lhs_layout(rank: int, all_gnodes, gnNo: int):
    """fsils_lhs_create's renumbering and overlap lists through the C ABI (b200_lhs_layout_*, csrc/lhs_layout.hpp; host side,
    needs no device): dict(map, mynNo, shnNo, reqs=[(peer, solver ids)]) for `rank` from every rank's global node list."""
    L = lib()
    lists = [_c(g, np.int32) for g in all_gnodes]
    counts = _c([len(g) for g in lists], np.int32)
    ptrs = (C.c_void_p * len(lists))(*[g.ctypes.data for g in lists])
    h = C.c_void_p()
    if L.b200_lhs_layout_create(int(rank), len(lists), int(gnNo), _p(counts), ptrs, C.byref(h)) != 0:
        raise RuntimeError("b200_lhs_layout_create: " + L.b200_last_error(None).decode())
    try:
        nNo, mynNo, shnNo, nReq = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.b200_lhs_layout_sizes(h, C.byref(nNo), C.byref(mynNo), C.byref(shnNo), C.byref(nReq))
        mp = np.zeros(nNo.value, np.int32)
        L.b200_lhs_layout_map(h, _p(mp))
        reqs = []
        for i in range(nReq.value):
            peer, n = C.c_int(), C.c_int()
            L.b200_lhs_layout_req(h, i, C.byref(peer), C.byref(n), None)
            ptr = np.zeros(n.value, np.int32)
            L.b200_lhs_layout_req(h, i, None, None, _p(ptr))
            reqs.append((peer.value, ptr))
        return dict(map=mp, mynNo=mynNo.value, shnNo=shnNo.value, reqs=reqs)
    finally:
        L.b200_lhs_layout_free(h)


class Backend:
    """One handle = one (equation x GPU), like one LinearAlgebra object per equation per MPI rank."""

    def __init__(self, device: int = 0):
        self.L = lib()
        self.h = C.c_void_p()
        if self.L.b200_create(C.byref(self.h), device) != 0:
            raise RuntimeError("b200_create: " + self.L.b200_last_error(None).decode())
        self.nNo = 0
        self.nnz = 0
        self.dof = 0

    def close(self):
        if self.h:
            self.L.b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: " + self.L.b200_last_error(self.h).decode())

    # -- communicator --------------------------------------------------------------------------
    def unique_id(self) -> np.ndarray:
        return unique_id()

    def comm_init(self, rank, nranks, uid):
        uid = _c(uid, np.uint8)
        self._ck(self.L.b200_comm_init(self.h, rank, nranks, _p(uid)), "b200_comm_init")

    def comm_transport(self):
        """'p2p: ...' (own kernels over peer-mapped windows), 'nccl: <why not p2p>' or 'none: single rank'."""
        return self.L.b200_comm_transport(self.h).decode()

    # -- structure -----------------------------------------------------------------------------
    def lhs_create(self, gnNo, rowPtr, colPtr, map=None, mynNo=None, reqs=(), nFaces=0):
        rowPtr = _c(rowPtr, np.int32); colPtr = _c(colPtr, np.int32)
        nNo = len(rowPtr) - 1
        mp = None if map is None else _c(map, np.int32)
        rr = _c([r[0] for r in reqs], np.int32)
        rn = _c([len(r[1]) for r in reqs], np.int32)
        rp = _c(np.concatenate([np.asarray(r[1], np.int32) for r in reqs]) if reqs else np.zeros(0, np.int32), np.int32)
        self._ck(self.L.b200_lhs_create(self.h, int(gnNo), nNo, int(nNo if mynNo is None else mynNo), len(colPtr),
                                        _p(rowPtr), _p(colPtr), _p(mp), len(reqs), _p(rr), _p(rn), _p(rp), nFaces),
                 "b200_lhs_create")
        self.nNo, self.nnz = nNo, len(colPtr)

    def face_set(self, faIn, glob, dof, bGrp, val=None, shared=False):
        glob = _c(glob, np.int32)
        v = None if val is None else _c(val, np.float64)
        self._ck(self.L.b200_face_set(self.h, faIn, len(glob), dof, bGrp, _p(glob), _p(v), int(shared)), "b200_face_set")

    # -- assembly --------------------------------------------------------------------------------
    def mesh_set(self, ien, x, qmTET4=-1.0):
        ien = _c(ien, np.int32); x = _c(x, np.float64)
        self._ck(self.L.b200_mesh_set(self.h, ien.shape[1], ien.shape[0], _p(ien), _p(x), qmTET4), "b200_mesh_set")

    def zero(self, dof):
        self._ck(self.L.b200_zero(self.h, dof), "b200_zero")
        self.dof = dof

    def state_set(self, tDof, Ag, Yg, Bf=None):
        """Ag/Yg/Bf: numpy arrays or raw host addresses (pinned buffers)."""
        if isinstance(Ag, np.ndarray):
            Ag = _c(Ag, np.float64); Yg = _c(Yg, np.float64)
            Bf = None if Bf is None else _c(Bf, np.float64)
        self._keep = (Ag, Yg, Bf)
        self._ck(self.L.b200_state_set(self.h, tDof, _p(Ag), _p(Yg), _p(Bf)), "b200_state_set")

    def assemble_fluid(self, props: FluidProps):
        self._ck(self.L.b200_assemble_fluid(self.h, C.byref(props)), "b200_assemble_fluid")

    def disp_set(self, tDof, Dg, Do=None):
        Dg = _c(Dg, np.float64)
        Do = None if Do is None else _c(Do, np.float64)
        self._ck(self.L.b200_disp_set(self.h, tDof, _p(Dg), _p(Do)), "b200_disp_set")

    def assemble_struct(self, props: StructProps):
        self._ck(self.L.b200_assemble_struct(self.h, C.byref(props)), "b200_assemble_struct")

    def prestress_set(self, pS0=None, pstEq=False):
        """com_mod.pS0 (nNo, 6) for the next struct assemblies (None: no prestress); pstEq: also accumulate pSn / pSa."""
        a = None if pS0 is None else _c(pS0, np.float64)
        self._ck(self.L.b200_prestress_set(self.h, _p(a), int(pstEq)), "b200_prestress_set")

    def prestress_get(self):
        pSn = np.zeros((self.nNo, 6)); pSa = np.zeros(self.nNo)
        self._ck(self.L.b200_prestress_get(self.h, _p(pSn), _p(pSa)), "b200_prestress_get")
        return pSn, pSa

    def assemble_lelas(self, props: LelasProps):
        self._ck(self.L.b200_assemble_lelas(self.h, C.byref(props)), "b200_assemble_lelas")

    def assemble_ustruct(self, props: UstructProps):
        self._ck(self.L.b200_assemble_ustruct(self.h, C.byref(props)), "b200_assemble_ustruct")

    def ustruct_r(self, amg, ami, s, Ad):
        Ad = _c(Ad, np.float64)
        self._ck(self.L.b200_ustruct_r(self.h, amg, ami, s, _p(Ad)), "b200_ustruct_r")

    def get_Kd(self):
        Kd = np.empty((self.nnz, 12))
        self._ck(self.L.b200_get_Kd(self.h, _p(Kd)), "b200_get_Kd")
        return Kd

    def mesh_fibers(self, fN):
        fN = _c(fN, np.float64)
        self._ck(self.L.b200_mesh_fibers(self.h, 2, _p(fN)), "b200_mesh_fibers")

    def mesh_domains(self, nDmn, elem_dmn):
        ed = _c(elem_dmn, np.int32)
        self._ck(self.L.b200_mesh_domains(self.h, nDmn, _p(ed)), "b200_mesh_domains")

    def assemble_fsi(self, kinds, fluid_props_list, struct_props_list):
        """kinds[d] in {0 fluid, 1 struct}; *_props_list[d] = props of domain d (None where not that kind)."""
        n = len(kinds)
        k = _c(kinds, np.int32)
        fl = (FluidProps * n)(*[p if p is not None else FluidProps() for p in fluid_props_list])
        st = (StructProps * n)(*[p if p is not None else StructProps() for p in struct_props_list])
        self._ck(self.L.b200_assemble_fsi(self.h, n, _p(k), fl, st), "b200_assemble_fsi")

    def assemble_elem(self, eqN, lK, lR):
        eqN = _c(eqN, np.int32); lK = _c(lK, np.float64); lR = _c(lR, np.float64)
        self._ck(self.L.b200_assemble_elem(self.h, len(eqN), _p(eqN), _p(lK), _p(lR)), "b200_assemble_elem")

    def get_R(self):
        R = np.empty((self.nNo, self.dof))
        self._ck(self.L.b200_get_R(self.h, _p(R)), "b200_get_R")
        return R

    def set_R(self, R):
        R = _c(R, np.float64)
        self.dof = R.shape[1]
        self._ck(self.L.b200_set_R(self.h, self.dof, _p(R)), "b200_set_R")

    def add_R(self, R):
        R = _c(R, np.float64)
        self._ck(self.L.b200_add_R(self.h, R.shape[1], _p(R)), "b200_add_R")

    def get_Val(self):
        V = np.empty((self.nnz, self.dof * self.dof))
        self._ck(self.L.b200_get_Val(self.h, _p(V)), "b200_get_Val")
        return V

    def set_Val(self, V):
        V = _c(V, np.float64)
        dof = int(round(V.shape[1] ** 0.5))
        self.dof = dof
        self._ck(self.L.b200_set_Val(self.h, dof, _p(V)), "b200_set_Val")

    def commu_R(self):
        self._ck(self.L.b200_commu_R(self.h), "b200_commu_R")

    # -- solve -----------------------------------------------------------------------------------
    def solve(self, ls_type, prec, RI, GM=None, CG=None, incL=None, res=None, out=None, fetch=True):
        """RI/GM/CG: (relTol, absTol, mItr, sD).  Returns (X or None, info dict).
        `out`: optional preallocated (nNo, dof) array or raw pinned address for the solution."""
        def tol(t):
            return None if t is None else Tol(t[0], t[1], int(t[2]), int(t[3]) if len(t) > 3 else 0)
        tRI, tGM, tCG = tol(RI), tol(GM), tol(CG)
        incL_a = None if incL is None else _c(incL, np.int32)
        res_a = None if res is None else _c(res, np.float64)
        X = None
        if fetch:
            X = out if out is not None else np.empty((self.nNo, self.dof))
        o = LsOut()
        self._ck(self.L.b200_solve(self.h, ls_type, prec,
                                   C.byref(tRI), C.byref(tGM) if tGM else None, C.byref(tCG) if tCG else None,
                                   _p(incL_a), _p(res_a), _p(X), C.byref(o)), "b200_solve")
        info = dict(RI=sub_out_dict(o.RI), GM=sub_out_dict(o.GM), CG=sub_out_dict(o.CG), Resm=o.Resm, Resc=o.Resc)
        return X, info

    def pattern(self, tnNo, meshes):
        """lhsa on the device: meshes = list of IEN arrays (nEl, eNoN).  Returns rowPtr (tnNo+1), colPtr (nnz)."""
        self._ck(self.L.b200_pattern_begin(self.h, tnNo), "b200_pattern_begin")
        for ien in meshes:
            ien = _c(ien, np.int32)
            self._ck(self.L.b200_pattern_add_mesh(self.h, ien.shape[1], ien.shape[0], _p(ien)), "b200_pattern_add_mesh")
        nnz = C.c_int(0)
        self._ck(self.L.b200_pattern_finish(self.h, C.byref(nnz)), "b200_pattern_finish")
        rowPtr = np.empty(tnNo + 1, np.int32); colPtr = np.empty(nnz.value, np.int32)
        self._ck(self.L.b200_pattern_get(self.h, _p(rowPtr), _p(colPtr)), "b200_pattern_get")
        return rowPtr, colPtr

    def assemble_fluid_dmn(self, props):
        arr = (FluidProps * len(props))(*props)
        self._ck(self.L.b200_assemble_fluid_dmn(self.h, len(props), arr), "b200_assemble_fluid_dmn")

    def assemble_struct_dmn(self, props):
        arr = (StructProps * len(props))(*props)
        self._ck(self.L.b200_assemble_struct_dmn(self.h, len(props), arr), "b200_assemble_struct_dmn")

    # -- boundary-face (Neumann) assembly (b_assem_neu_bc) ---------------------------------------------
    def face_mesh_set(self, faIn, IENb, gE):
        IENb = _c(IENb, np.int32); gE = _c(gE, np.int32)
        self._ck(self.L.b200_face_mesh_set(self.h, faIn, IENb.shape[1], IENb.shape[0], _p(IENb), _p(gE)), "b200_face_mesh_set")

    def assemble_bfolw(self, faIn, hg, *, dt, af, beta=0.0, tDof=3, s=0, ustruct=False, am=1.0, gam=0.0):
        """b_neu_folw_p: follower pressure load on a struct (dof 3) or ustruct (dof 4, + Kd) face."""
        p = BfolwProps()
        p.dt, p.af, p.beta, p.tDof, p.s = dt, af, beta, tDof, s
        p.ustruct, p.am, p.gam = int(ustruct), am, gam
        hg = _c(hg, np.float64)
        self._ck(self.L.b200_assemble_bfolw(self.h, faIn, C.byref(p), _p(hg)), "b200_assemble_bfolw")

    def face_integ(self, faIn, which, l=0, u=None, geo=0):
        """all_fun::integ over face faIn of rows l..u of a device array ("Yn", "Yo", ... or None for the area)."""
        out = C.c_double(0.0)
        u = l if u is None else u
        self._ck(self.L.b200_face_integ(self.h, faIn, -1 if which is None else PIC[which], l, u, geo, C.byref(out)), "b200_face_integ")
        return out.value

    def face_normal_update(self, faIn, lsFace, geo=2):
        """fsi_ls_upd: recompute the face vector of linear-solver face lsFace from face mesh faIn (geo as in face_integ)."""
        self._ck(self.L.b200_face_normal_update(self.h, faIn, lsFace, geo), "b200_face_normal_update")

    def face_get_val(self, lsFace, nNoFace, dof=3):
        val = np.empty((nNoFace, dof))
        self._ck(self.L.b200_face_get_val(self.h, lsFace, _p(val)), "b200_face_get_val")
        return val

    def assemble_bneu(self, faIn, kind, hg, *, dt=0.0, af=0.0, gam=0.0, tDof=4, mvMsh=False, rho=0.0, bfs=0.0):
        """kind "fluid" (b_fluid) or "solid" (b_l_elas); hg: nodal Neumann values (nNo,)."""
        p = BneuProps()
        p.dt, p.af, p.gam, p.tDof, p.mvMsh, p.rho, p.bfs = dt, af, gam, tDof, int(mvMsh), rho, bfs
        hg = _c(hg, np.float64)
        self._ck(self.L.b200_assemble_bneu(self.h, faIn, {"fluid": 0, "solid": 1}[kind], C.byref(p), _p(hg)), "b200_assemble_bneu")

    # -- time integrator on the device (pic::picp / pici / picc) ---------------------------------------
    def pic_init(self, tDof, eqs, dFlag=False, sstEq=False):
        """eqs: list of dict(s, e, am, af, gam, beta, kind)."""
        arr = (PicEq * len(eqs))()
        for i, q in enumerate(eqs):
            arr[i].s, arr[i].e, arr[i].kind = int(q["s"]), int(q["e"]), int(q.get("kind", 0))
            arr[i].am, arr[i].af, arr[i].gam, arr[i].beta = q["am"], q["af"], q["gam"], q.get("beta", 0.0)
        self._ck(self.L.b200_pic_init(self.h, tDof, len(eqs), arr, int(dFlag), int(sstEq)), "b200_pic_init")
        self.pic_tDof = tDof

    def pic_set(self, which, a):
        a = _c(a, np.float64)
        self._ck(self.L.b200_pic_set(self.h, PIC[which], _p(a)), "b200_pic_set")

    def pic_get(self, which):
        a = np.empty((self.nNo, 3 if which == "Ad" else self.pic_tDof))
        self._ck(self.L.b200_pic_get(self.h, PIC[which], _p(a)), "b200_pic_get")
        return a

    def pic_scatter(self, which, idx, val):
        idx = _c(idx, np.int32); val = _c(val, np.float64)
        self._ck(self.L.b200_pic_scatter(self.h, PIC[which], len(idx), _p(idx), _p(val)), "b200_pic_scatter")

    def picp(self, dt):
        self._ck(self.L.b200_picp(self.h, dt), "b200_picp")

    def pici(self):
        self._ck(self.L.b200_pici(self.h), "b200_pici")

    def picc(self, iEq, dt, first_itr=False):
        self._ck(self.L.b200_picc(self.h, iEq, dt, int(first_itr)), "b200_picc")

    def pic_copy_rows(self, nodes, s2, cnt):
        nodes = _c(nodes, np.int32)
        self._ck(self.L.b200_pic_copy_rows(self.h, len(nodes), _p(nodes), s2, cnt), "b200_pic_copy_rows")

    def pic_advance(self):
        self._ck(self.L.b200_pic_advance(self.h), "b200_pic_advance")

    # -- taps --------------------------------------------------------------------------------------
    def spmv(self, x):
        x = _c(x, np.float64)
        y = np.empty_like(x)
        self._ck(self.L.b200_spmv(self.h, x.shape[1], _p(x), _p(y)), "b200_spmv")
        return y

    def op_bench(self, op, k=0, reps=20):
        """Stand-alone bench of one kernel class (name from KERNEL_CLASSES or id): (ms, bytes) per launch."""
        if isinstance(op, str):
            op = KERNEL_CLASSES.index(op)
        ms = C.c_double(0); by = C.c_double(0)
        self._ck(self.L.b200_op_bench(self.h, op, k, reps, C.byref(ms), C.byref(by)), "b200_op_bench")
        return ms.value, by.value

    def fp64_peak(self, k=16, reps=5):
        """FP64 FMA peak of the device in TFLOP/s (k_fma_peak: 8 independent DFMA chains per thread, 148 x 8 CTAs)."""
        ms, flops = self.op_bench(100, k=k, reps=reps)
        return flops / (ms * 1e-3) / 1e12

    def tune(self, name, value):
        """Kernel-variant knob (b200_tune): 'vv3', 'schur_gp', 'schur_sp', 'narrow', 'cg_batch'."""
        self.L.b200_tune.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        self._ck(self.L.b200_tune(self.h, name.encode(), int(value)), "b200_tune")

    def profile(self, enable=True):
        self._ck(self.L.b200_profile(self.h, int(enable)), "b200_profile")

    def profile_read(self):
        n = len(KERNEL_CLASSES)
        ms = np.zeros(n); by = np.zeros(n); ln = np.zeros(n, np.int64)
        self._ck(self.L.b200_profile_read(self.h, n, _p(ms), _p(by), _p(ln)), "b200_profile_read")
        return {k: dict(ms=float(ms[i]), bytes=float(by[i]), launches=int(ln[i])) for i, k in enumerate(KERNEL_CLASSES)}

    def timer_start(self):
        self._ck(self.L.b200_timer(self.h, 0, None), "b200_timer")

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        self._ck(self.L.b200_timer(self.h, 1, C.byref(ms)), "b200_timer")
        return ms.value

    def launch_count(self) -> int:
        return int(self.L.b200_launch_count(self.h))

    def timings(self):
        t = np.zeros(4)
        self.L.b200_last_timings(self.h, _p(t))
        return dict(assembly_ms=t[0], precond_ms=t[1], krylov_ms=t[2])
