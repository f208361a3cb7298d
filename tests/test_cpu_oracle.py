"""CPU suite (`-m "not gpu"`): the oracle against the golden fixtures, the synthetic-mesh generator
against the reference's own lhsa, the product's solver control flow (krylov.hpp with the test-only
host policy) against the compiled reference, and the C-ABI export check."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import have_ref, needs_ref
from util import ROOT, golden, hl_solve, rel_inf, rel_l2

from svfsiplus_b200 import mesh as M
from svfsiplus_b200 import backend as B
from svfsiplus_b200 import problem as P


def test_mesh_counts_match_survey():
    # nx=ny=8,nz=16 -> 6144 tets; Kuhn split: nnz = nNo + 2*edges
    m = M.pipe_mesh(8, 8, 16)
    assert m.nEl == 6144 and m.nNo == 9 * 9 * 17
    assert (M.tet_volumes(m.x, m.ien) > 0).all()
    rowPtr, colPtr = M.csr_pattern(m.ien, m.nNo)
    e = np.sort(m.ien[:, [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]].reshape(-1, 2), axis=1)
    nedges = len(np.unique(e[:, 0].astype(np.int64) * m.nNo + e[:, 1]))
    assert len(colPtr) == m.nNo + 2 * nedges
    # sorted columns, diagonal present
    for a in (0, 17, m.nNo - 1):
        row = colPtr[rowPtr[a]:rowPtr[a + 1]]
        assert (np.diff(row) > 0).all() and a in row


def test_p10_sizes_formula():
    # P10 = 96x96x181 hexes: 10 008 576 tets, 1 712 438 nodes (checked by formula, not generated here)
    nx, ny, nz = 96, 96, 181
    assert 6 * nx * ny * nz == 10_008_576
    assert (nx + 1) * (ny + 1) * (nz + 1) == 1_712_438


@needs_ref
def test_csr_pattern_equals_reference_lhsa():
    from oracle import ref
    m = M.pipe_mesh(5, 4, 7)
    ra = ref.RefAssembly(m.x, m.ien)
    rp, cp = ra.csr()
    rp2, cp2 = M.csr_pattern(m.ien, m.nNo)
    assert (rp == rp2).all() and (cp == cp2).all()
    w, N, Nx = ra.tables()
    s = (5.0 + 3.0 * np.sqrt(5.0)) / 20.0
    t = (1.0 - s) / 3.0
    assert np.allclose(w, 1.0 / 24.0)
    assert N[0, 0] == s and N[0, 1] == t and N[3, 3] == 1.0 - t - t - t


@needs_ref
def test_oracle_reproduces_golden_fixtures():
    """The compiled reference, re-run now, reproduces the committed fixtures bit for bit."""
    from oracle import refcase
    g = golden("pipe_4_4_6.npz")
    case = P.pipe_case(4, 4, 6)
    R, Val, rowPtr, colPtr, _ = refcase.reference_assemble(case)
    assert (rowPtr == g["rowPtr"]).all() and (colPtr == g["colPtr"]).all()
    assert np.array_equal(R, g["R"]) and np.array_equal(Val, g["Val"])
    for ls in ("NS", "GMRES", "CG", "BICGS"):
        X, o = refcase.reference_solve(case, R, Val, P.LS_SETTINGS[ls])
        assert np.array_equal(X, g[f"X_{ls}"]), ls
        assert int(o["itr"]) == int(g[f"info_{ls}"][1])


def _ls_vec(ls):
    from oracle import refcase
    return refcase._ls_vector(P.LS_SETTINGS[ls])


@pytest.mark.parametrize("ls", ["NS", "GMRES", "CG", "BICGS"])
@pytest.mark.parametrize("coupled", [False, True])
def test_hostlogic_matches_golden(ls, coupled):
    """krylov.hpp control flow == reference fsils_solve (iteration counts equal, solution <= 1e-8)."""
    g = golden("pipe_4_4_6.npz")
    case = P.pipe_case(4, 4, 6, coupled=coupled)
    if coupled:
        X, V, o = hl_solve(g["rowPtr"], g["colPtr"], 4, g["R"], g["Val"], _ls_vec(ls), 701, case["faces"], case["incL"], case["res"])
        assert rel_l2(X, g[f"X_{ls}"]) < 1e-8
        assert abs(int(o["itr"]) - int(g[f"info_{ls}"][1])) <= 1
        assert abs(int(o["GM_itr"]) - int(g[f"info_{ls}"][4])) <= 2
    else:
        # uncoupled outlet: no fixture; must still converge consistently with the oracle when present
        X, V, o = hl_solve(g["rowPtr"], g["colPtr"], 4, g["R"], g["Val"], _ls_vec(ls), 701, case["faces"], case["incL"], case["res"])
        assert np.isfinite(X).all()


@needs_ref
@pytest.mark.parametrize("ls", ["NS", "GMRES", "CG", "BICGS"])
def test_hostlogic_matches_reference_mid_mesh(ls):
    from oracle import refcase
    case = P.pipe_case(8, 8, 16)
    R, Val, rowPtr, colPtr, _ = refcase.reference_assemble(case)
    Xr, oref = refcase.reference_solve(case, R, Val, P.LS_SETTINGS[ls])
    X, V, o = hl_solve(rowPtr, colPtr, 4, R, Val, _ls_vec(ls), 701, case["faces"], case["incL"], case["res"])
    assert rel_l2(X, Xr) < 1e-8
    assert int(o["itr"]) == int(oref["itr"])
    assert int(o["GM_itr"]) == int(oref["GM_itr"]) and int(o["CG_itr"]) == int(oref["CG_itr"])
    assert abs(o["fNorm"] - oref["fNorm"]) <= 1e-6 * abs(oref["fNorm"])


def test_hostlogic_res_required_error():
    g = golden("pipe_4_4_6.npz")
    case = P.pipe_case(4, 4, 6)
    with pytest.raises(RuntimeError, match="res is required for Neu surfaces"):
        hl_solve(g["rowPtr"], g["colPtr"], 4, g["R"], g["Val"], _ls_vec("GMRES"), 701, case["faces"], case["incL"], None)


def test_c_abi_exports_every_declared_symbol():
    """libsvb200.so loads on a CPU-only box and exports every function include/svb200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "svb200.h")).read()
    declared = sorted(set(re.findall(r"\b(b200_[A-Za-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(os.path.join(ROOT, "svfsiplus_b200", "libsvb200.so"))
    for name in declared:
        assert hasattr(lib, name), name
    from svfsiplus_b200 import backend
    assert sorted(backend.EXPORTS) == declared


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product refuses to create a handle (no silent CPU path)."""
    from svfsiplus_b200 import backend as B
    if B.lib().b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        B.Backend(0)


@pytest.mark.parametrize("elem,eNoN", [("tet", 4), ("hex", 8)])
def test_gauss_tables_equal_reference(elem, eNoN):
    """The element kernels' Gauss rule / shape tables (host code of the product, no device needed) against
    what the reference's select_ele leaves in lM.w / lM.N / lM.Nx (stored in the golden fixture)."""
    from svfsiplus_b200 import backend as B
    g = golden("block_3_solid.npz")
    w, N, Nxi = B.elem_tables(eNoN)
    assert np.array_equal(w, g[f"w_{elem}"])
    assert np.abs(N - g[f"N_{elem}"]).max() < 1e-15
    assert np.abs(Nxi - g[f"Nx_{elem}"]).max() < 1e-15


def test_oracle_reproduces_solid_golden():
    if not have_ref():
        pytest.skip("compiled reference not available")
    from oracle import refcase
    g = golden("block_3_solid.npz")
    for elem in ("tet", "hex"):
        for kind, iso, vol in (("struct", "nHook", "ST91"), ("lelas", None, None), ("mesh", None, None)):
            c = P.block_case(3, elem=elem, kind=kind, iso=iso or "nHook", vol=vol)
            R, Val, *_ = refcase.reference_assemble_solid(c)
            tag = f"{elem}_{kind}_{iso}_{vol}"
            assert np.array_equal(R, g[f"R_{tag}"]) and np.array_equal(Val, g[f"Val_{tag}"])


# ---------------------------------------------------------------------------------------------------------------------------
# edge cases of the solver control flow (csrc/krylov.hpp, the header the device instantiates with CudaOps) against fsils_solve
# ---------------------------------------------------------------------------------------------------------------------------
EDGE_LS = {
    # restarts: Krylov dimension 5, up to 40 outer iterations
    "gmres_restarts": (B.LS_GMRES, (1e-6, 1e-17, 40, 5), None, None),
    # iteration limit reached before the tolerance: suc = false, the iterate so far is returned
    "gmres_not_converged": (B.LS_GMRES, (1e-14, 1e-30, 1, 3), None, None),
    "cg_not_converged": (B.LS_CG, (1e-14, 1e-30, 3, 0), None, None),
    "bicgs_not_converged": (B.LS_BICGS, (1e-14, 1e-30, 2, 0), None, None),
    # absolute tolerance above the initial norm: immediate return
    "gmres_abstol": (B.LS_GMRES, (1e-3, 1e30, 4, 50), None, None),
    "bicgs_abstol": (B.LS_BICGS, (1e-3, 1e30, 50, 0), None, None),
    "cg_abstol": (B.LS_CG, (1e-3, 1e30, 50, 0), None, None),
    # NS block solver: one outer iteration, tight and loose inner solves, tiny inner Krylov spaces (the last two make the
    # reference give up: its exception text must come out of the product's control flow as well)
    "ns_one_outer": (B.LS_NS, (1e-3, 1e-17, 1, 250), (1e-3, 1e-17, 10, 250), (1e-3, 1e-17, 300, 0)),
    "ns_tight_inner": (B.LS_NS, (1e-5, 1e-17, 8, 250), (1e-6, 1e-17, 20, 30), (1e-6, 1e-17, 500, 0)),
    "ns_small_inner_space": (B.LS_NS, (1e-3, 1e-17, 10, 250), (1e-2, 1e-17, 3, 4), (1e-1, 1e-17, 5, 0)),
    "ns_outer_space_of_five": (B.LS_NS, (1e-6, 1e-17, 40, 5), None, None),
}


def _both(case, R, Val, ls):
    """(X, info) or the exception text, from the reference and from the product's control flow on the host policy."""
    from oracle import refcase
    g = golden("pipe_4_4_6.npz")
    out = []
    for who in ("ref", "host"):
        try:
            if who == "ref":
                X, o = refcase.reference_solve(case, R, Val, ls)
            else:
                X, _, o = hl_solve(g["rowPtr"], g["colPtr"], 4, R, Val, refcase._ls_vector(ls), 701, case["faces"], case["incL"], case["res"])
            out.append((X, o, None))
        except RuntimeError as e:
            out.append((None, None, str(e)))
    return out


@needs_ref
@pytest.mark.parametrize("name", sorted(EDGE_LS))
def test_hostlogic_edge_cases_match_reference(name):
    """Same success flag and iteration counters (+-1) as the reference, same solution (<= 1e-8) - or the same exception text -
    for restarts, iteration limits, absolute-tolerance exits and the NS solver's inner limits."""
    g = golden("pipe_4_4_6.npz")
    case = P.pipe_case(4, 4, 6)
    ls = EDGE_LS[name]
    (Xr, oref, er), (X, o, eh) = _both(case, g["R"], g["Val"], ls)
    assert er == eh, (er, eh)
    if er is not None:
        assert er.startswith("FSILS:")
        return
    assert bool(o["suc"]) == bool(oref["suc"])
    assert abs(int(o["itr"]) - int(oref["itr"])) <= 1
    if ls[0] == B.LS_NS:
        assert abs(int(o["GM_itr"]) - int(oref["GM_itr"])) <= 2 and abs(int(o["CG_itr"]) - int(oref["CG_itr"])) <= 3
    scale = max(np.linalg.norm(Xr), 1e-300)
    assert np.linalg.norm(X - Xr) / scale < 1e-8
    assert np.isfinite(X).all()


@needs_ref
@pytest.mark.parametrize("ls", ["NS", "GMRES", "CG", "BICGS"])
def test_hostlogic_zero_right_hand_side(ls):
    """R = 0: whatever the reference does (zero vector without iterating, NaNs from the zero norm, or 'Singular matrix detected'
    from the NS solver's Gram system) the product's control flow does too."""
    g = golden("pipe_4_4_6.npz")
    case = P.pipe_case(4, 4, 6)
    R0 = np.zeros_like(g["R"])
    (Xr, oref, er), (X, o, eh) = _both(case, R0, g["Val"], P.LS_SETTINGS[ls])
    assert er == eh, (er, eh)
    if er is not None:
        return
    assert np.array_equal(np.isnan(X), np.isnan(Xr))
    assert np.array_equal(np.nan_to_num(X), np.nan_to_num(Xr))
    assert int(o["itr"]) == int(oref["itr"])


@needs_ref
@pytest.mark.parametrize("elem", ["tet", "hex"])
@pytest.mark.parametrize("ls", ["BICGS_STRUCT", "GMRES_STRUCT", "GMRES_STRUCT_LOOSE", "CG_MESH"])
def test_hostlogic_dof3_systems_match_reference(elem, ls):
    """The 3-dof block systems (struct / mesh equation: spar_mul_vv with dof 3, Dirichlet faces only) through the product's
    control flow against fsils_solve: same counters, solution <= 1e-8."""
    from oracle import refcase
    kind = "mesh" if ls == "CG_MESH" else "struct"
    case = P.block_case(3, elem=elem, kind=kind)
    R, Val, rowPtr, colPtr, *_ = refcase.reference_assemble_solid(case)
    m = case["mesh"]
    part = dict(gnNo=m.nNo, gNodes=np.arange(m.nNo), rowPtr=rowPtr, colPtr=colPtr,
                faces=[dict(nodes=f["nodes"], dof=f["dof"], bGrp=f["bGrp"], val=f["val"]) for f in case["faces"]])
    from oracle import ref
    rr = ref.RefRanks([part])
    Xr, _, oref = rr.solve(3, _ls_vec(ls), ref.PREC_FSILS, [R], [Val], case["incL"], case["res"])
    rr.close()
    X, V, o = hl_solve(rowPtr, colPtr, 3, R, Val, _ls_vec(ls), 701, case["faces"], case["incL"], case["res"])
    assert bool(o["suc"]) == bool(oref[0]["suc"])
    assert abs(int(o["itr"]) - int(oref[0]["itr"])) <= 1
    assert rel_l2(X, Xr[0]) < 1e-8


def test_c_client_of_the_hot_path_builds_and_refuses_to_run_without_a_device(tmp_path):
    """examples/newton_step.c - the C ABI's call sequence for one Newton iteration in plain C99 - compiles and links against
    libsvb200.so; on a box without a GPU it reports that and exits with status 3 (no CPU fallback), with one it solves."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    libdir = os.path.join(ROOT, "svfsiplus_b200")
    exe = tmp_path / "newton_step"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "newton_step.c"),
                    "-L" + libdir, "-lsvb200", "-Wl,-rpath," + libdir, "-lm", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    from svfsiplus_b200 import backend as B
    if B.lib().b200_device_count() > 0:
        assert r.returncode == 0 and "GMRES:" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr
