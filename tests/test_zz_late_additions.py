"""GPU parity of the features added late in round 1 (HO-ma law, solid viscosity for struct and ustruct), developed on the CPU
through the host/device-shared headers and the compiled reference and confirmed on a B200 with the round's last GPU seconds.
The file name sorts last on purpose: the driver runs `pytest -x`, and these were the least exercised tests of the round.

Tolerances as everywhere (BASELINE.json north_star): assembled R / Val / Kd <= 1e-12 relative (max-norm)."""
import numpy as np
import pytest

from util import golden, rel_inf

from svfsiplus_b200 import problem as P

# First run on a B200 at the very end of round 1: 16 passed (profiles/r01_late_additions_gpu_tests.log).
pytestmark = pytest.mark.gpu

TOL_ASM = 1e-12


@pytest.mark.parametrize("elem", ["tet", "hex"])
def test_struct_ho_ma_matches_golden(elem):
    """stIso_HO_ma in struct_3d (mat_models_carray.h:1137-1353)."""
    g = golden("late_additions.npz")
    case = P.block_case(3, elem=elem, kind="struct", iso="HO_ma", vol="ST91")
    be = P.setup_backend(case)
    P.assemble_solid(be, case)
    assert rel_inf(be.get_R(), g[f"R_{elem}_struct_HO_ma"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{elem}_struct_HO_ma"]) < TOL_ASM
    be.close()


@pytest.mark.parametrize("elem", ["tet", "hex"])
def test_ustruct_ho_ma_matches_golden(elem):
    """stIso_HO_ma in ustruct_3d_m / _c (get_pk2cc_dev, mat_models.cpp:963-1054)."""
    g = golden("late_additions.npz")
    case = P.ustruct_case(3, elem=elem, iso="HO_ma")
    be = P.setup_backend(case)
    P.assemble_ustruct(be, case)
    assert rel_inf(be.get_R(), g[f"R_{elem}_ustruct_HO_ma"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{elem}_ustruct_HO_ma"]) < TOL_ASM
    assert rel_inf(be.get_Kd(), g[f"Kd_{elem}_ustruct_HO_ma"]) < TOL_ASM
    be.close()


@pytest.mark.parametrize("visc", ["newt", "pot"])
@pytest.mark.parametrize("elem,n", [("tet", 3), ("hex", 3), ("tet10", 2)])
def test_struct_solid_viscosity_matches_golden(elem, n, visc):
    """dmn.solid_visc (get_visc_stress_and_tangent<3>, mat_models_carray.h:1578) in struct_3d: the viscous stress in the
    residual and afu*Kvis_u + afv*Kvis_v in the tangent (sv_struct.cpp:666-675, 771-842)."""
    g = golden("late_additions.npz")
    case = P.block_case(n, elem=elem, kind="struct", iso="nHook", vol="ST91", visc=visc, visc_mu=5.0e4)
    be = P.setup_backend(case)
    P.assemble_solid(be, case)
    R, Val = be.get_R(), be.get_Val()
    assert rel_inf(R, g[f"R_{elem}_struct_visc_{visc}"]) < TOL_ASM
    assert rel_inf(Val, g[f"Val_{elem}_struct_visc_{visc}"]) < TOL_ASM
    # the inviscid kernel is a different instantiation: the viscous terms must be visible
    g0 = golden("block_3_solid.npz") if n == 3 else None
    if g0 is not None:
        assert rel_inf(Val, g0[f"Val_{elem}_struct_nHook_ST91"]) > 1e-2
    P.assemble_solid(be, case, upload=False)
    assert np.array_equal(R, be.get_R()) and np.array_equal(Val, be.get_Val())
    be.close()


@pytest.mark.parametrize("visc", ["newt", "pot"])
@pytest.mark.parametrize("elem,n", [("tet", 3), ("hex", 3), ("tet10", 2)])
def test_ustruct_solid_viscosity_matches_golden(elem, n, visc):
    """dmn.solid_visc in ustruct_3d_m (ustruct.cpp:1275-1302, 1406-1550): Siso + Svis, Kvis_u in Ku (lK and lKd), af Kvis_v."""
    g = golden("late_additions.npz")
    case = P.ustruct_case(n, elem=elem, visc=visc, visc_mu=5.0e4)
    be = P.setup_backend(case)
    P.assemble_ustruct(be, case)
    assert rel_inf(be.get_R(), g[f"R_{elem}_ustruct_visc_{visc}"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{elem}_ustruct_visc_{visc}"]) < TOL_ASM
    assert rel_inf(be.get_Kd(), g[f"Kd_{elem}_ustruct_visc_{visc}"]) < TOL_ASM
    be.close()


@pytest.mark.parametrize("visc", [None, "pot"])
@pytest.mark.parametrize("elem,n", [("tet", 3), ("hex", 3), ("tet10", 2)])
def test_struct_prestress_matches_golden(elem, n, visc):
    """com_mod.pS0 / pstEq in construct_dsolid + struct_3d (sv_struct.cpp:262-345, 646-700): S += S0 interpolated from the nodal
    prestress, and the accumulations pSn(:,A) += w N_a pSl, pSa(A) += w N_a, with and without solid viscosity."""
    g = golden("late_additions.npz")
    case = P.block_case(n, elem=elem, kind="struct", iso="nHook", vol="ST91", visc=visc, visc_mu=5.0e4, prestress=True)
    be = P.setup_backend(case)
    P.assemble_solid(be, case)
    tag = f"{elem}_struct_pst_{visc}"
    assert rel_inf(be.get_R(), g[f"R_{tag}"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{tag}"]) < TOL_ASM
    pSn, pSa = be.prestress_get()
    assert rel_inf(pSn, g[f"pSn_{tag}"]) < TOL_ASM
    assert rel_inf(pSa, g[f"pSa_{tag}"]) < TOL_ASM
    # switching the prestress off again gives the plain element back
    be.prestress_set(None, False)
    P.assemble_solid(be, dict(case, pS0=None, pstEq=False), upload=False)
    g0 = golden("late_additions.npz")[f"R_{elem}_struct_visc_pot"] if visc else (golden("block_3_solid.npz")[f"R_{elem}_struct_nHook_ST91"] if n == 3 else None)
    if g0 is not None:
        assert rel_inf(be.get_R(), g0) < TOL_ASM
    with pytest.raises(RuntimeError, match="prestress_get"):
        be.prestress_get()
    be.close()


# ---- end-to-end and FSI-wall cases (goldens from the compiled reference) ------------------------------------------------------


@pytest.mark.parametrize("tag", ["tet", "hex", "tet10"])
def test_fsi_with_prestressed_viscous_wall_matches_golden(tag):
    """construct_fsi with com_mod.pS0 (read, never accumulated: fsi.cpp:147-148, 225) and dmn.solid_visc on the struct domain: the
    extended struct element writing into the dof-4 blocks."""
    g = golden("late_additions.npz")
    case = {"tet": lambda: P.fsi_case(4, 4, 4), "hex": lambda: P.fsi_block_case(3, elem="hex"), "tet10": lambda: P.fsi_block_case(2, elem="tet10")}[tag]()
    rng = np.random.default_rng(77)
    case["solid"] = dict(case["solid"], visc="pot", visc_mu=200.0)
    case["pS0"] = 1.0e4 * rng.standard_normal((case["mesh"].nNo, 6))
    be = P.setup_backend(case)
    P.assemble_fsi(be, case)
    assert rel_inf(be.get_R(), g[f"R_{tag}_fsi_wall"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{tag}_fsi_wall"]) < TOL_ASM
    be.close()


def test_time_step_result_file_passes_the_reference_harness_criterion(tmp_path):
    """One Newton-converged time step with the state resident on the device, written as result_001.vtu by the VTK-free writer, against
    the same step made of the reference's own functions written the same way: compared file against file with the reference test
    harness's own per-field criterion (tests/conftest.py: Velocity 1e-7, Pressure 1e-6; sv_io.compare_results)."""
    from oracle import ref, refcase
    from svfsiplus_b200 import sv_io as IO
    if not ref.available():
        pytest.skip("oracle/_ref not present on this box")
    case = P.pipe_case(8, 8, 16, coupled=False)
    p, m = case["props"], case["mesh"]
    dt, n_newton = p["dt"], 6
    eqs = [dict(s=0, e=3, am=p["am"], af=p["af"], gam=p["gam"], beta=0.0, phys="fluid", kind=0)]
    zeros = np.zeros((m.nNo, 4))
    be = P.setup_backend(case)
    be.pic_init(4, eqs)
    be.pic_set("Ao", case["Ag"]); be.pic_set("Yo", case["Yg"]); be.pic_set("Do", zeros)
    be.picp(dt)
    be.state_set(4, None, None, case["Bf"])
    for it in range(n_newton):
        be.pici()
        P.newton_linear_step(be, case, ls="NS", upload=False, fetch=False)
        be.picc(0, dt, first_itr=(it == 0))
    Yn_g = be.pic_get("Yn")
    be.close()
    st = dict(Ao=case["Ag"], Yo=case["Yg"], Do=zeros, An=zeros, Yn=zeros, Dn=zeros, Ad=np.zeros((m.nNo, 3)), Ag=zeros, Yg=zeros, Dg=zeros)
    st = ref.pic("p", st, eqs, dt=dt)
    for it in range(n_newton):
        st = ref.pic("i", st, eqs, dt=dt)
        c = dict(case); c["Ag"] = st["Ag"]; c["Yg"] = st["Yg"]
        _, _, X, _ = refcase.reference_step(c, "NS")
        st = ref.pic("c", st, eqs, dt=dt, R=X, Rd=np.zeros((m.nNo, 3)))
    os_ = __import__("os")
    os_.makedirs(tmp_path / "b200"); os_.makedirs(tmp_path / "ref")
    mine = IO.write_results(str(tmp_path / "b200" / "result"), 1, m.x, m.ien, 10, {"Velocity": Yn_g[:, :3], "Pressure": Yn_g[:, 3]})
    theirs = IO.write_results(str(tmp_path / "ref" / "result"), 1, m.x, m.ien, 10, {"Velocity": st["Yn"][:, :3], "Pressure": st["Yn"][:, 3]})
    assert IO.compare_results(mine, theirs, ["Velocity", "Pressure"]) == []


@pytest.mark.parametrize("elem", ["hex", "tet"])
def test_solid_block_two_time_steps_match_the_complete_reference(elem):
    """End to end against the reference ITSELF: tests/golden/full_reference_runs.npz holds Displacement / Velocity after two time steps
    of the exported solid block (tools/export_case.py --block 3; struct, neo-Hookean + ST91, traction 5e6 on Z1, one displacement
    component held on X0 / Y0 / Z0, BICG 1e-12, three Newton iterations per step) computed by the complete reference solver
    (oracle/_ref/svmultiphysics_ref).  Here the same two steps run with the state resident on the device: picp -> [pici -> struct
    assembly + Neumann face -> BICG -> picc] x 3 -> advance, compared with the reference harness's criterion (|a - b| <= rtol + rtol |b|)."""
    import os
    from svfsiplus_b200 import backend as B
    from svfsiplus_b200 import mesh as M
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "full_reference_runs.npz"))
    case = P.block_case(3, elem=elem, kind="struct", iso="nHook", vol="ST91")
    m, p = case["mesh"], case["props"]
    # solver.xml: X0 holds the z component, Y0 the y component, Z0 the x component (Effective_direction); Z1 carries the traction
    faces = []
    for nm, comp in (("X0", 2), ("Y0", 1), ("Z0", 0)):
        nodes = m.faces[nm]["nodes"]
        val = np.ones((len(nodes), 3)); val[:, comp] = 0.0
        faces.append(dict(name=nm, nodes=nodes, dof=3, bGrp=B.BC_DIR, val=val))
    z1 = m.faces["Z1"]["nodes"]
    faces.append(dict(name="Z1", nodes=z1, dof=3, bGrp=B.BC_NEU, val=np.zeros((len(z1), 3))))
    case = dict(case, faces=faces, incL=np.array([1, 1, 1, 0], np.int32), res=np.zeros(4))
    be = P.setup_backend(case)
    on = np.zeros(m.nNo, bool); on[z1] = True
    IENb, gE = M.face_elements(m, on)
    be.face_mesh_set(3, IENb, gE)
    hg = np.zeros(m.nNo); hg[z1] = -5.0e6                       # set_bc_neu_l: hg = -g * gx (set_bc.cpp:1431-1434)
    dt = p["dt"]
    eqs = [dict(s=0, e=2, am=p["am"], af=p["af"], gam=p["gam"], beta=p["beta"], phys="struct", kind=0)]
    zeros = np.zeros((m.nNo, 3))
    be.pic_init(3, eqs, dFlag=True)
    be.pic_set("Ao", zeros); be.pic_set("Yo", zeros); be.pic_set("Do", zeros)
    be.state_set(3, None, None, zeros)
    case0 = dict(case, Ag=zeros, Yg=zeros, Dg=zeros, Bf=zeros)
    ls_type, RI, GM, CG = P.LS_SETTINGS["BICGS_STRUCT"]
    for step in range(2):
        be.picp(dt)
        for it in range(3):
            be.pici()
            P.assemble_solid(be, case0, upload=False)
            be.assemble_bneu(3, "solid", hg, tDof=3, dt=dt, af=p["af"], gam=p["gam"])
            be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case["incL"], case["res"], fetch=False)
            be.picc(0, dt, first_itr=(it == 0))
        be.pic_advance()
    Dn, Yn = be.pic_get("Do"), be.pic_get("Yo")               # after the advance the new state is the old one of the next step
    be.close()
    for got, name, rtol in ((Dn, "Displacement", 1.0e-10), (Yn, "Velocity", 1.0e-7)):
        want = g[f"block_{elem}_3/{name}"]
        assert (np.abs(got - want) - rtol - rtol * np.abs(want) <= 0.0).all(), name
