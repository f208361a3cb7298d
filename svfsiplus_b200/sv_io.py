"""ctypes front-end of the VTK-free I/O library (include/svb200_io.h -> svfsiplus_b200/libsvb200io.so, host/sv_io.cpp).

Mirrors what the reference does with VtkData / vtk_xml.cpp / output.cpp (file:line in the header): read a .vtu / .vtp mesh,
write a result file, write / read restart records, format history lines.  Test and tooling front-end only: numpy in, numpy
out; every byte of format logic is in the C++ library."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

POINT_DATA, CELL_DATA = 0, 1
ASCII, BINARY, APPENDED_RAW, APPENDED_BASE64 = 0, 1, 2, 3
VTK_TYPE = {"LINE": 3, "TRI3": 5, "QUD4": 9, "TET4": 10, "HEX8": 12, "WEDGE": 13, "TRI6": 22, "QUD8": 23, "TET10": 24, "HEX20": 25,
            "QUD9": 28, "HEX27": 29}


class Restart(C.Structure):
    _fields_ = [("stamp", C.c_int * 7), ("cTS", C.c_int), ("time", C.c_double), ("cpu_time", C.c_double),
                ("nEq", C.c_int), ("iNorm", C.c_void_p), ("nXn", C.c_int), ("xn", C.c_void_p),
                ("tDof", C.c_int), ("tnNo", C.c_int), ("Yn", C.c_void_p), ("An", C.c_void_p),
                ("dFlag", C.c_int), ("Dn", C.c_void_p), ("trailing_Dn", C.c_int),
                ("sstEq", C.c_int), ("nsd", C.c_int), ("Ad", C.c_void_p),
                ("pstEq", C.c_int), ("nsymd", C.c_int), ("pS0", C.c_void_p)]


class History(C.Structure):
    _fields_ = [("sym", C.c_char_p), ("cTS", C.c_int), ("itr", C.c_int), ("saved", C.c_int),
                ("elapsed", C.c_double), ("since_last", C.c_double), ("eq_iNorm", C.c_double), ("eq_pNorm", C.c_double),
                ("ri_iNorm", C.c_double), ("ri_fNorm", C.c_double), ("ri_dB", C.c_double), ("ri_callD", C.c_double),
                ("ri_itr", C.c_int), ("ri_suc", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsvb200io.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python __graft_entry__.py` (build()) first")
        L = C.CDLL(path)
        L.b200io_last_error.restype = C.c_char_p
        L.b200io_vtk_new.restype = C.c_void_p
        L.b200io_vtk_array_name.restype = C.c_char_p
        L.b200io_restart_record_bytes.restype = C.c_longlong
        for name, args in {
            "b200io_vtk_read": [C.c_char_p, C.c_void_p], "b200io_vtk_is_polydata": [C.c_void_p], "b200io_vtk_num_points": [C.c_void_p],
            "b200io_vtk_num_cells": [C.c_void_p], "b200io_vtk_nodes_per_cell": [C.c_void_p], "b200io_vtk_points": [C.c_void_p, C.c_void_p],
            "b200io_vtk_connectivity": [C.c_void_p, C.c_void_p], "b200io_vtk_cell_types": [C.c_void_p, C.c_void_p],
            "b200io_vtk_num_arrays": [C.c_void_p, C.c_int], "b200io_vtk_array_name": [C.c_void_p, C.c_int, C.c_int],
            "b200io_vtk_array_info": [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "b200io_vtk_array_f64": [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p], "b200io_vtk_array_i32": [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p],
            "b200io_vtk_free": [C.c_void_p], "b200io_vtk_new": [C.c_int], "b200io_vtk_set_points": [C.c_void_p, C.c_int, C.c_void_p],
            "b200io_vtk_set_cells": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int],
            "b200io_vtk_add_array_f64": [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p],
            "b200io_vtk_add_array_i32": [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p],
            "b200io_vtk_write": [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int],
            "b200io_restart_record_bytes": [C.c_void_p], "b200io_restart_write": [C.c_char_p, C.c_int, C.c_longlong, C.c_void_p, C.c_int],
            "b200io_restart_read": [C.c_char_p, C.c_int, C.c_longlong, C.c_void_p], "b200io_restart_name": [C.c_char_p, C.c_int, C.c_char_p, C.c_int],
            "b200io_history_header": [C.c_int, C.c_char_p, C.c_int], "b200io_history_line": [C.c_void_p, C.c_char_p, C.c_int],
        }.items():
            getattr(L, name).argtypes = args
        _LIB = L
    return _LIB


class IoError(RuntimeError):
    pass


def _ck(rc):
    if rc != 0:
        raise IoError(lib().b200io_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def read_vtk(path):
    """read_vtu / read_vtp (vtk_xml.cpp:568, :438): dict(x (nNo,3), ien (nEl,eNoN) or the ragged pair (conn, offsets), types, point_data, cell_data)."""
    L = lib()
    h = C.c_void_p()
    _ck(L.b200io_vtk_read(os.fsencode(path), C.byref(h)))
    try:
        nNo, nEl, eNoN = L.b200io_vtk_num_points(h), L.b200io_vtk_num_cells(h), L.b200io_vtk_nodes_per_cell(h)
        out = dict(polydata=bool(L.b200io_vtk_is_polydata(h)), nNo=nNo, nEl=nEl, eNoN=eNoN)
        x = np.zeros((nNo, 3))
        _ck(L.b200io_vtk_points(h, _p(x)))
        out["x"] = x
        if eNoN > 0:
            ien = np.zeros((nEl, eNoN), np.int32)
            _ck(L.b200io_vtk_connectivity(h, _p(ien)))
            out["ien"] = ien
        types = np.zeros(nEl, np.uint8)
        if nEl:
            _ck(L.b200io_vtk_cell_types(h, _p(types)))
        out["types"] = types
        for where, key in ((POINT_DATA, "point_data"), (CELL_DATA, "cell_data")):
            d = {}
            for i in range(L.b200io_vtk_num_arrays(h, where)):
                name = L.b200io_vtk_array_name(h, where, i)
                nc, nt, isint = C.c_int(), C.c_int(), C.c_int()
                assert L.b200io_vtk_array_info(h, where, name, C.byref(nc), C.byref(nt), C.byref(isint)) == 0
                a = np.zeros((nt.value, nc.value), np.int32 if isint.value else np.float64)
                _ck((L.b200io_vtk_array_i32 if isint.value else L.b200io_vtk_array_f64)(h, where, name, _p(a)))
                d[name.decode()] = a[:, 0] if nc.value == 1 else a
            out[key] = d
        return out
    finally:
        L.b200io_vtk_free(h)


def write_vtk(path, x, ien, vtk_type, point_data=None, cell_data=None, *, polydata=False, mode=APPENDED_RAW, compress=True, header64=True):
    """write_vtu / write_vtp / write_vtus (vtk_xml.cpp:855, :827, :913): one element type, named point / cell arrays."""
    L = lib()
    x = np.ascontiguousarray(x, np.float64)
    ien = np.ascontiguousarray(ien, np.int32)
    h = C.c_void_p(L.b200io_vtk_new(int(polydata)))
    try:
        _ck(L.b200io_vtk_set_points(h, x.shape[0], _p(x)))
        _ck(L.b200io_vtk_set_cells(h, ien.shape[0], ien.shape[1] if ien.ndim == 2 else 1, _p(ien), int(vtk_type)))
        for where, d in ((POINT_DATA, point_data or {}), (CELL_DATA, cell_data or {})):
            for name, a in d.items():
                a = np.asarray(a)
                nt = a.shape[0]
                nc = 1 if a.ndim == 1 else a.shape[1]
                if np.issubdtype(a.dtype, np.integer):
                    a = np.ascontiguousarray(a, np.int32)
                    _ck(L.b200io_vtk_add_array_i32(h, where, name.encode(), nc, nt, _p(a)))
                else:
                    a = np.ascontiguousarray(a, np.float64)
                    _ck(L.b200io_vtk_add_array_f64(h, where, name.encode(), nc, nt, _p(a)))
        _ck(L.b200io_vtk_write(h, os.fsencode(path), mode, int(compress), int(header64)))
    finally:
        L.b200io_vtk_free(h)


def _restart_struct(keep, *, stamp, cTS, time, cpu_time, iNorm, xn, Yn, An, Dn=None, Ad=None, pS0=None, trailing_Dn=True):
    r = Restart()
    r.stamp[:] = [int(v) for v in stamp]
    r.cTS, r.time, r.cpu_time = int(cTS), float(time), float(cpu_time)

    def put(a):
        a = np.ascontiguousarray(a, np.float64)
        keep.append(a)
        return a, a.ctypes.data

    a, r.iNorm = put(iNorm); r.nEq = a.size
    a, r.xn = put(xn); r.nXn = a.size
    a, r.Yn = put(Yn); r.tnNo, r.tDof = a.shape
    a, r.An = put(An)
    r.dFlag = int(Dn is not None)
    if Dn is not None:
        _, r.Dn = put(Dn)
    r.trailing_Dn = int(trailing_Dn)
    r.sstEq = int(Ad is not None)
    if Ad is not None:
        a, r.Ad = put(Ad); r.nsd = a.shape[1]
    r.pstEq = int(pS0 is not None)
    if pS0 is not None:
        a, r.pS0 = put(pS0); r.nsymd = a.shape[1]
    return r


def restart_record_bytes(**kw):
    keep = []
    r = _restart_struct(keep, **kw)
    return int(lib().b200io_restart_record_bytes(C.byref(r)))


def write_restart(path, rank, recLn, create=True, **kw):
    """output::write_restart (output.cpp:202-345).  Arrays are (tnNo, tDof) row-major = the solver's (tDof, tnNo) column-major."""
    keep = []
    r = _restart_struct(keep, **kw)
    _ck(lib().b200io_restart_write(os.fsencode(path), rank, recLn, C.byref(r), int(create)))


def read_restart(path, rank, recLn, *, nEq, nXn, tDof, tnNo, dFlag=False, nsd=0, nsymd=0):
    """init_from_bin (initialize.cpp:81-230): returns dict(stamp, cTS, time, cpu_time, iNorm, xn, Yn, An[, Dn][, Ad][, pS0])."""
    kw = dict(stamp=[0]*7, cTS=0, time=0.0, cpu_time=0.0, iNorm=np.zeros(nEq), xn=np.zeros(nXn), Yn=np.zeros((tnNo, tDof)), An=np.zeros((tnNo, tDof)))
    if dFlag:
        kw["Dn"] = np.zeros((tnNo, tDof))
    if nsd:
        kw["Ad"] = np.zeros((tnNo, nsd))
    if nsymd:
        kw["pS0"] = np.zeros((tnNo, nsymd))
    keep = []
    r = _restart_struct(keep, **kw)
    _ck(lib().b200io_restart_read(os.fsencode(path), rank, recLn, C.byref(r)))
    names = ["iNorm", "xn", "Yn", "An"] + (["Dn"] if dFlag else []) + (["Ad"] if nsd else []) + (["pS0"] if nsymd else [])
    out = dict(zip(names, keep))
    out.update(stamp=list(r.stamp), cTS=r.cTS, time=r.time, cpu_time=r.cpu_time)
    return out


def restart_name(stem, cTS):
    buf = C.create_string_buffer(4096)
    lib().b200io_restart_name(os.fsencode(stem), cTS, buf, len(buf))
    return buf.value.decode()


def history_header(nEq):
    buf = C.create_string_buffer(1024)
    lib().b200io_history_header(nEq, buf, len(buf))
    return buf.value.decode()


def history_line(sym, cTS, itr, *, saved=False, elapsed, since_last, eq_iNorm, eq_pNorm, ri_iNorm, ri_fNorm, ri_dB, ri_callD, ri_itr, ri_suc):
    h = History(sym.encode(), cTS, itr, int(saved), elapsed, since_last, eq_iNorm, eq_pNorm, ri_iNorm, ri_fNorm, ri_dB, ri_callD, ri_itr, int(ri_suc))
    buf = C.create_string_buffer(1024)
    lib().b200io_history_line(C.byref(h), buf, len(buf))
    return buf.value.decode()


# ---------------------------------------------------------------------------------------------------------------------------
# result comparison with the reference's own acceptance criterion
# ---------------------------------------------------------------------------------------------------------------------------
# per-field relative tolerances of the reference's test harness (tests/conftest.py:18-35)
RTOL = {
    "Action_potential": 1.0e-10, "Cauchy_stress": 1.0e-4, "Concentration": 1.0e-10, "Def_grad": 1.0e-10, "Divergence": 1.0e-9,
    "Displacement": 1.0e-10, "Jacobian": 1.0e-10, "Pressure": 1.0e-6, "Stress": 1.0e-4, "Strain": 1.0e-10, "Temperature": 1.0e-10,
    "Traction": 1.0e-6, "Velocity": 1.0e-7, "VonMises_stress": 1.0e-3, "Vorticity": 1.0e-7, "WSS": 1.0e-8,
}


def compare_results(result_vtu, reference_vtu, fields, rtol=None):
    """run_with_reference's check (tests/conftest.py:150-200) on two result files read with the VTK-free reader: every point-data
    field must satisfy |a - b| <= rtol + rtol |b| entry by entry (rtol doubles as the absolute floor, as in the reference).
    A 2-D result against a 3-D reference drops the reference's zero third component.  Returns a list of failure messages
    (empty = pass); raises ValueError for a missing field or a field without a tolerance, like the reference."""
    res, ref = read_vtk(result_vtu), read_vtk(reference_vtu)
    tol = dict(RTOL, **(rtol or {}))
    msgs = []
    for f in fields:
        if f not in res["point_data"]:
            raise ValueError("Field " + f + " not in simulation result")
        if f not in ref["point_data"]:
            raise ValueError("Field " + f + " not in reference result")
        if f not in tol:
            raise ValueError("No tolerance defined for field " + f)
        a, b = np.asarray(res["point_data"][f], np.float64), np.asarray(ref["point_data"][f], np.float64)
        if a.ndim == 2 and b.ndim == 2 and a.shape[1] == 2 and b.shape[1] == 3:
            assert not np.any(b[:, 2])
            b = b[:, :2]
        if a.shape != b.shape:
            msgs.append(f"Test failed in field {f}. Shapes differ: {a.shape} vs {b.shape}")
            continue
        r = tol[f]
        a_fl, b_fl = a.ravel(), b.ravel()
        rel_diff = np.abs(a_fl - b_fl) - r - r * np.abs(b_fl)
        close = rel_diff <= 0.0
        if not np.all(close):
            i = int(rel_diff.argmax())
            msgs.append(f"Test failed in field {f}. Results differ by more than rtol={r} in {1 - close.sum() / close.size:.1%} of results. "
                        f"Max. rel. difference is {rel_diff[i]:.1e} (abs. {abs(a_fl[i] - b_fl[i]):.1e})")
    return msgs


def result_name(save_name, cTS):
    """write_vtus' file name (vtk_xml.cpp:1344-1352): <saveName>_<cTS, three digits up to 1000>.vtu"""
    return f"{save_name}_{cTS}.vtu" if cTS > 1000 else f"{save_name}_{cTS:03d}.vtu"


def write_results(save_name, cTS, x, ien, vtk_type, fields, domain_id=None, proc_id=None, **kw):
    """The file write_vtus leaves for one mesh (vtk_xml.cpp:913-1470): the named nodal output fields as point data, in the order
    given (a dict name -> (nNo,) or (nNo, ncomp); the reference's names are Velocity, Pressure, Displacement, WSS, ...), Domain_ID /
    Proc_ID as cell data when given.  Returns the path; compare two such files with compare_results."""
    path = result_name(save_name, cTS)
    cd = {}
    if domain_id is not None:
        cd["Domain_ID"] = np.asarray(domain_id, np.int32)
    if proc_id is not None:
        cd["Proc_ID"] = np.asarray(proc_id, np.int32)
    write_vtk(path, x, ien, vtk_type, {k: np.asarray(v, np.float64) for k, v in fields.items()}, cd, **kw)
    return path
