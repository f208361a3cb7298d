"""Constitutive laws of the solid elements (csrc/solid_law.hpp, the source the device kernels compile) on the host against
the compiled reference's mat_models_carray::get_pk2cc<3> (Code/Source/solver/mat_models_carray.h:182) at random deformation
gradients: 2nd Piola-Kirchhoff stress and the Voigt elasticity matrix, every law x penalty x fibre stress combination."""
import numpy as np
import pytest

from conftest import needs_ref
from util import host_pk2cc, host_visc

from svfsiplus_b200 import mesh as M

E = 240.56596e6
MU = 0.5 * E / 1.5
HO = dict(a=590.0, b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12, afs=2160.0, bfs=11.436, khs=100.0)
LAWS = [
    ("nHook", "ST91", dict(C10=0.5 * MU, Kpen=4.0e9)),
    ("nHook", "M94", dict(C10=0.5 * MU, Kpen=4.0e9, Tfa=3.0e4, eta_s=0.4)),
    ("nHook", "Quad", dict(C10=0.5 * MU, Kpen=1.0e8)),
    ("StVK", None, dict(C10=E * 0.3 / (1.3 * 0.4), C01=0.5 * E / 1.3)),
    ("mStVK", None, dict(C10=E / 1.2, C01=0.5 * E / 1.3)),
    ("HO", "ST91", dict(Kpen=1.0e6, ho=HO)),
    ("HO", "M94", dict(Kpen=1.0e6, ho=HO, Tfa=2.0e4, eta_s=0.3)),
    ("MR", "ST91", dict(C10=0.3 * MU, C01=0.2 * MU, Kpen=4.0e9)),
    ("MR", None, dict(C10=0.3 * MU, C01=0.2 * MU)),
    ("MR", "M94", dict(C10=0.3 * MU, C01=0.2 * MU, Kpen=1.0e8, Tfa=3.0e4, eta_s=0.4)),
    ("HGO", "ST91", dict(C10=0.5 * MU, Kpen=4.0e9, kap=0.226, ho=dict(aff=9.96e5, bff=524.6, ass=9.96e5, bss=524.6))),
    ("Gucci", "ST91", dict(C10=880.0, Kpen=1.0e6, ho=dict(bff=8.0, bss=6.0, bfs=12.0))),
    ("Gucci", "M94", dict(C10=2000.0, Kpen=1.0e5, ho=dict(bff=18.5, bss=3.58, bfs=1.63), Tfa=500.0, eta_s=0.4)),
    ("HO_ma", "ST91", dict(Kpen=1.0e6, ho=HO)),
    ("HO_ma", "M94", dict(Kpen=1.0e6, ho=HO, Tfa=2.0e4, eta_s=0.3)),
    ("HO_ma", "Quad", dict(Kpen=1.0e5, ho=dict(HO, khs=20.0))),
    ("HGO", "Quad", dict(C10=0.5 * MU, Kpen=1.0e8, kap=0.0, ho=dict(aff=2.0e6, bff=20.0, ass=1.0e6, bss=10.0), Tfa=3.0e4, eta_s=0.4)),
]
VOIGT = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (2, 0)]


@pytest.mark.parametrize("iso,vol,kw", LAWS, ids=[f"{l[0]}-{l[1]}-{'Tf' if 'Tfa' in l[2] else '0'}" for l in LAWS])
@needs_ref
def test_law_matches_reference_get_pk2cc(iso, vol, kw):
    from oracle import ref
    m = M.block_mesh(1, "tet")
    ra = ref.RefAssembly(m.x, m.ien)
    rng = np.random.default_rng(42)
    th = 0.7
    fl = np.array([np.cos(th), np.sin(th), 0.0, -np.sin(th), np.cos(th), 0.0])
    worst = 0.0
    for _ in range(20):
        F = np.eye(3) + 0.15 * rng.standard_normal((3, 3))
        if np.linalg.det(F) < 0.3:
            continue
        S, Dm = ra.pk2cc(F, fl, iso=iso, vol=vol, **kw)
        S6, Dm21 = host_pk2cc(F, fl, iso=iso, vol=vol, **kw)
        Sr = np.array([S[i, j] for i, j in VOIGT])
        Dr = np.array([Dm[I, J] for I in range(6) for J in range(I, 6)])
        worst = max(worst, np.abs(S6 - Sr).max() / np.abs(Sr).max(), np.abs(Dm21 - Dr).max() / np.abs(Dr).max())
        assert np.allclose(Dm, Dm.T, rtol=1e-10, atol=1e-10 * np.abs(Dm).max())          # the reference's Dm is symmetric
    ra.close()
    # the isochoric projections are evaluated in closed form here and by generic fourth-order contractions there
    assert worst < 1e-12, worst


DEV_LAWS = [l for l in LAWS if l[0] not in ("StVK", "mStVK")]


@pytest.mark.parametrize("iso,vol,kw", DEV_LAWS, ids=[f"{l[0]}-{'Tf' if 'Tfa' in l[2] else '0'}-{i}" for i, l in enumerate(DEV_LAWS)])
@needs_ref
def test_law_matches_reference_get_pk2cc_dev(iso, vol, kw):
    """The ustruct kernels evaluate the same law with Kpen = 0: it must equal mat_models::get_pk2cc_dev (mat_models.cpp:630)."""
    from oracle import ref
    m = M.block_mesh(1, "tet")
    ra = ref.RefAssembly(m.x, m.ien)
    rng = np.random.default_rng(43)
    th = 0.4
    fl = np.array([np.cos(th), 0.0, np.sin(th), 0.0, 1.0, 0.0])
    kw = dict(kw, Kpen=0.0)
    worst = 0.0
    for _ in range(12):
        F = np.eye(3) + 0.15 * rng.standard_normal((3, 3))
        if np.linalg.det(F) < 0.3:
            continue
        S, Dm = ra.pk2cc(F, fl, iso=iso, vol=None, dev=True, **kw)
        S6, Dm21 = host_pk2cc(F, fl, iso=iso, vol=None, **kw)
        Sr = np.array([S[i, j] for i, j in VOIGT])
        Dr = np.array([Dm[I, J] for I in range(6) for J in range(I, 6)])
        worst = max(worst, np.abs(S6 - Sr).max() / np.abs(Sr).max(), np.abs(Dm21 - Dr).max() / np.abs(Dr).max())
    ra.close()
    assert worst < 1e-12, worst


@needs_ref
def test_oracle_reproduces_late_addition_fixtures():
    """tests/golden/late_additions.npz (make_golden_late.py) is what the compiled reference gives today."""
    from oracle import refcase
    from util import golden
    from svfsiplus_b200 import problem as P
    g = golden("late_additions.npz")
    for elem in ("tet", "hex"):
        c = P.block_case(3, elem=elem, kind="struct", iso="HO_ma", vol="ST91")
        R, Val, *_ = refcase.reference_assemble_solid(c)
        assert np.array_equal(R, g[f"R_{elem}_struct_HO_ma"]) and np.array_equal(Val, g[f"Val_{elem}_struct_HO_ma"])
        c = P.ustruct_case(3, elem=elem, iso="HO_ma")
        R, Val, Kd, _ = refcase.reference_assemble_ustruct(c)
        assert np.array_equal(R, g[f"R_{elem}_ustruct_HO_ma"]) and np.array_equal(Val, g[f"Val_{elem}_ustruct_HO_ma"])
        assert np.array_equal(Kd, g[f"Kd_{elem}_ustruct_HO_ma"])
    for tag, c in (("tet", P.fsi_case(4, 4, 4)), ("hex", P.fsi_block_case(3, elem="hex")), ("tet10", P.fsi_block_case(2, elem="tet10"))):
        rng = np.random.default_rng(77)
        c["solid"] = dict(c["solid"], visc="pot", visc_mu=200.0)
        c["pS0"] = 1.0e4 * rng.standard_normal((c["mesh"].nNo, 6))
        R, Val, _ = refcase.reference_assemble_fsi(c)
        assert np.array_equal(R, g[f"R_{tag}_fsi_wall"]) and np.array_equal(Val, g[f"Val_{tag}_fsi_wall"])
    for elem, n in (("tet", 3), ("hex", 3), ("tet10", 2)):
        for visc in ("newt", "pot"):
            c = P.block_case(n, elem=elem, kind="struct", iso="nHook", vol="ST91", visc=visc, visc_mu=5.0e4)
            R, Val, *_ = refcase.reference_assemble_solid(c)
            assert np.array_equal(R, g[f"R_{elem}_struct_visc_{visc}"]) and np.array_equal(Val, g[f"Val_{elem}_struct_visc_{visc}"])
            for v2 in ((None, "pot") if visc == "pot" else ()):
                c = P.block_case(n, elem=elem, kind="struct", iso="nHook", vol="ST91", visc=v2, visc_mu=5.0e4, prestress=True)
                R, Val, *_ = refcase.reference_assemble_solid(c)
                tag = f"{elem}_struct_pst_{v2}"
                assert np.array_equal(R, g[f"R_{tag}"]) and np.array_equal(Val, g[f"Val_{tag}"])
                assert np.array_equal(c["_ref_pSn"], g[f"pSn_{tag}"]) and np.array_equal(c["_ref_pSa"], g[f"pSa_{tag}"])
            c = P.ustruct_case(n, elem=elem, visc=visc, visc_mu=5.0e4)
            R, Val, Kd, _ = refcase.reference_assemble_ustruct(c)
            assert np.array_equal(R, g[f"R_{elem}_ustruct_visc_{visc}"]) and np.array_equal(Val, g[f"Val_{elem}_ustruct_visc_{visc}"])
            assert np.array_equal(Kd, g[f"Kd_{elem}_ustruct_visc_{visc}"])


@needs_ref
@pytest.mark.parametrize("model", ["newt", "pot"])
@pytest.mark.parametrize("eNoN", [4, 8, 10])
def test_solid_viscosity_matches_reference_bitwise(model, eNoN):
    """visc_point / visc_pair (solid_law.hpp) against get_visc_stress_and_tangent<3> (mat_models_carray.h:1578): the viscous
    stress and both tangent arrays of every node pair, bit for bit (same sums in the same order, no FMA contraction)."""
    from oracle import ref
    rng = np.random.default_rng(99 + eNoN)
    for _ in range(6):
        F = np.eye(3) + 0.2 * rng.standard_normal((3, 3))
        if np.linalg.det(F) < 0.3:
            continue
        vx = 5.0 * rng.standard_normal((3, 3))
        Nx = rng.standard_normal((eNoN, 3))
        Sr, Kur, Kvr = ref.visc(model, 50.0, Nx, vx, F)
        S, Ku, Kv = host_visc(model, 50.0, Nx, vx, F)
        assert np.array_equal(S, Sr)
        assert np.array_equal(Ku, Kur) and np.array_equal(Kv, Kvr)


@needs_ref
@pytest.mark.parametrize("visc,prestress", [("newt", False), ("pot", False), (None, True), ("pot", True)])
@pytest.mark.parametrize("elem", ["tet", "hex"])
def test_struct_extended_element_restated_in_numpy_matches_reference(elem, visc, prestress):
    """The arithmetic k_assemble_solid<..., VISC = true> performs (dv/dX per Gauss point, S = S_el + Svis, pSl = S, S += S0 from the
    nodal prestress, P = F S, the residual row, T1 + afu (BtDB + Kvis_u) + afv Kvis_v per node pair, and the pstEq sums
    pSn += w N_a pSl, pSa += w N_a), restated element by element in numpy on top of the SAME
    host/device-shared point functions (pk2cc_iso, visc_point, visc_pair), against construct_dsolid of the compiled reference.
    Pins the way the viscous terms enter the element (index conventions, afu / afv, pair order) without a GPU."""
    from oracle import refcase
    from svfsiplus_b200 import backend as B
    from svfsiplus_b200 import problem as P
    c = P.block_case(2, elem=elem, kind="struct", iso="nHook", vol="ST91", visc=visc, visc_mu=5.0e4, prestress=prestress)
    Rr, Vr, rowPtr, colPtr, *_ = refcase.reference_assemble_solid(c)
    m, pr = c["mesh"], c["props"]
    eNoN = m.ien.shape[1]
    w, N, Nxi = B.elem_tables(eNoN)
    dt, am, af, gam, beta, rho, dmp = (pr[k] for k in ("dt", "am", "af", "gam", "beta", "rho", "dmp"))
    afu, afv, amd = af * beta * dt * dt, af * gam * dt, am * rho + af * gam * dt * dmp
    R = np.zeros_like(Rr)
    Val = np.zeros_like(Vr)
    pSn = np.zeros((m.nNo, 6)); pSa = np.zeros(m.nNo)
    pos = {}
    for A in range(m.nNo):
        for p in range(rowPtr[A], rowPtr[A + 1]):
            pos[(A, colPtr[p])] = p
    VO = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (2, 0)]
    for e in range(m.nEl):
        nd = m.ien[e]
        xl, al, yl, dl, bl = m.x[nd], c["Ag"][nd], c["Yg"][nd], c["Dg"][nd], c["Bf"][nd]
        for g in range(len(w)):
            xXi = xl.T @ Nxi[g]                                   # dx/dxi
            Jac = np.linalg.det(xXi)
            Nx = Nxi[g] @ np.linalg.inv(xXi)                     # (a, 3): dN/dX
            F = np.eye(3) + dl.T @ Nx
            vx = yl.T @ Nx
            ud = -rho * np.asarray(pr["f"]) + N[g] @ (rho * (al - bl) + dmp * yl)
            S6, Dm21 = host_pk2cc(F, np.zeros(6), iso="nHook", vol="ST91", C10=pr["C10"], Kpen=pr["Kpen"])
            if visc:
                Sv, Ku, Kv = host_visc(visc, pr["visc_mu"], Nx, vx, F)
            else:
                Sv, Ku, Kv = np.zeros((3, 3)), np.zeros((eNoN, eNoN, 3, 3)), np.zeros((eNoN, eNoN, 3, 3))
            S = np.zeros((3, 3)); Dm = np.zeros((6, 6))
            for k, (i, j) in enumerate(VO):
                S[i, j] = S[j, i] = S6[k] + Sv[i, j]             # the six entries the kernel keeps
            wj = w[g] * Jac
            if prestress:
                pSl = np.array([S[i, j] for i, j in VO])
                pSn[nd] += wj * np.outer(N[g], pSl)
                pSa[nd] += wj * N[g]
                S0 = N[g] @ c["pS0"][nd]
                for k, (i, j) in enumerate(VO):
                    S[i, j] = S[j, i] = S[i, j] + S0[k]
            it = iter(Dm21)
            for I in range(6):
                for J in range(I, 6):
                    Dm[I, J] = Dm[J, I] = next(it)
            Pk = F @ S
            wj = w[g] * Jac
            Bm = np.zeros((eNoN, 6, 3))
            for a in range(eNoN):
                for k, (i, j) in enumerate(VO):
                    Bm[a, k] = Nx[a, i] * F[:, i] if i == j else Nx[a, i] * F[:, j] + F[:, i] * Nx[a, j]
            for a in range(eNoN):
                R[nd[a]] += wj * (N[g, a] * ud + Pk @ Nx[a])
                for b in range(eNoN):
                    T1 = amd * N[g, a] * N[g, b] + afu * (Nx[a] @ S @ Nx[b])
                    K = wj * (T1 * np.eye(3) + afu * (Bm[a].T @ Dm @ Bm[b] + Ku[a, b]) + afv * Kv[a, b])
                    Val[pos[(nd[a], nd[b])]] += K.reshape(9)
    assert np.abs(R - Rr).max() / np.abs(Rr).max() < 1e-12
    assert np.abs(Val - Vr).max() / np.abs(Vr).max() < 1e-12
    if prestress:
        assert np.abs(pSn - c["_ref_pSn"]).max() / np.abs(c["_ref_pSn"]).max() < 1e-12
        assert np.abs(pSa - c["_ref_pSa"]).max() / np.abs(c["_ref_pSa"]).max() < 1e-12


@needs_ref
@pytest.mark.parametrize("visc", ["newt", "pot"])
@pytest.mark.parametrize("elem", ["tet", "hex"])
def test_ustruct_viscosity_terms_restated_in_numpy_match_reference(elem, visc):
    """What k_assemble_ustruct<..., VISC = true> adds to the inviscid element (ustruct.cpp:1275-1302, 1406-1550):
        Siso += Svis   ->  dR_a = w F Svis Nx_a,   dKd_ab = w af (Kvis_u(a,b) + (Nx_a . Svis Nx_b) I),
        dK_ab  = w af Kvis_v(a,b) + (af/am) dKd_ab          (af = eq.af eq.gam dt)
    rebuilt in numpy from the host/device-shared visc_point / visc_pair, against the DIFFERENCE of two runs of the compiled
    reference's construct_usolid (with and without dmn.solid_visc)."""
    from oracle import refcase
    from svfsiplus_b200 import backend as B
    from svfsiplus_b200 import problem as P
    c1 = P.ustruct_case(2, elem=elem, visc=visc, visc_mu=5.0e4)
    c0 = P.ustruct_case(2, elem=elem)
    R1, V1, K1, _ = refcase.reference_assemble_ustruct(c1)
    R0, V0, K0, _ = refcase.reference_assemble_ustruct(c0)
    m, pr = c1["mesh"], c1["props"]
    rowPtr, colPtr = c1["rowPtr"], c1["colPtr"]
    eNoN = m.ien.shape[1]
    w, N, Nxi = B.elem_tables(eNoN)
    af = pr["af"] * pr["gam"] * pr["dt"]
    afm = af / pr["am"]
    dR = np.zeros_like(R0); dV = np.zeros_like(V0); dK = np.zeros_like(K0)
    pos = {(A, colPtr[p]): p for A in range(m.nNo) for p in range(rowPtr[A], rowPtr[A + 1])}
    for e in range(m.nEl):
        nd = m.ien[e]
        xl, yl, dl = m.x[nd], c1["Yg"][nd][:, :3], c1["Dg"][nd][:, :3]
        for g in range(len(w)):
            xXi = xl.T @ Nxi[g]
            Nx = Nxi[g] @ np.linalg.inv(xXi)
            wj = w[g] * np.linalg.det(xXi)
            F = np.eye(3) + dl.T @ Nx
            vx = yl.T @ Nx
            Sv, Ku, Kv = host_visc(visc, pr["visc_mu"], Nx, vx, F)
            Sv = 0.5 * (Sv + Sv.T)
            Pv = F @ Sv
            for a in range(eNoN):
                dR[nd[a], :3] += wj * (Pv @ Nx[a])
                for b in range(eNoN):
                    kd = wj * af * (Ku[a, b] + (Nx[a] @ Sv @ Nx[b]) * np.eye(3))
                    p = pos[(nd[a], nd[b])]
                    dK[p].reshape(4, 3)[:3] += kd
                    dV[p].reshape(4, 4)[:3, :3] += wj * af * Kv[a, b] + afm * kd
    for got, want, ref0 in ((dR, R1 - R0, R1), (dV, V1 - V0, V1), (dK, K1 - K0, K1)):
        assert np.abs(got - want).max() / np.abs(ref0).max() < 1e-12
        assert np.abs(want).max() > 0
