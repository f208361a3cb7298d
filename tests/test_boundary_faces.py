"""Boundary-face (Neumann) assembly (SURVEY.md par. 8(f) row 1): b_assem_neu_bc + gnnb + b_fluid / b_l_elas
(Code/Source/solver/eq_assem.cpp:58, nn.cpp:552, fluid.cpp:46, l_elas.cpp:48) on TRI3 / QUD4 / TRI6 faces.

CPU: the host/device-shared arithmetic (csrc/face_elem.hpp) against the compiled reference, bit for bit.
GPU: b200_face_mesh_set + b200_assemble_bneu through the C ABI against the reference at 1e-12, alone and on top of a
volume assembly, with a moving mesh, and the size-independent property sum_a R_a = -area * n for a unit traction."""
import numpy as np
import pytest

from conftest import needs_ref
from util import golden, host_bfolw_assemble, host_bneu_assemble, host_face_integ, host_face_normals, rel_inf

from svfsiplus_b200 import mesh as M
from svfsiplus_b200 import problem as P

TIME = dict(dt=0.005, af=0.6, gam=0.7)
ELEMS = [("tet", 3), ("hex", 3), ("tet10", 2)]


def _setup(elem, n, mvMsh=False, seed=0):
    case = P.fluid_block_case(n, elem=elem, mvMsh=mvMsh)
    m = case["mesh"]
    on = np.abs(m.x[:, 2]) < 1e-12                             # Z0: the flow enters there, so the backflow term is active
    IENb, gE = M.face_elements(m, on)
    rng = np.random.default_rng(seed)
    hg = np.where(on, 50.0 + 10.0 * rng.standard_normal(m.nNo), 0.0)
    Do = None
    if mvMsh:
        Do = np.zeros((m.nNo, 7))
        Do[:, 4:7] = 0.03 / n * rng.standard_normal((m.nNo, 3))
    return case, IENb, gE, hg, Do


def _ref_face(case, kind, IENb, gE, hg, Do, mvMsh):
    from oracle import ref
    m = case["mesh"]
    ra = ref.RefAssembly(m.x, m.ien)
    Yg = case["Yg"] if kind == "fluid" else case["Yg"][:, :3]
    R, Val = ra.bneu(kind, IENb, gE, hg, Yg, rho=1.06, bfs=0.2, mvMsh=mvMsh, Do=Do, **TIME)
    ra.close()
    return R, Val


# ---------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("elem,n", ELEMS)
@pytest.mark.parametrize("kind,mvMsh", [("fluid", False), ("fluid", True), ("solid", False)])
@needs_ref
def test_host_face_element_matches_reference_bitwise(elem, n, kind, mvMsh):
    case, IENb, gE, hg, Do = _setup(elem, n, mvMsh)
    Rr, Vr = _ref_face(case, kind, IENb, gE, hg, Do, mvMsh)
    Yg = case["Yg"] if kind == "fluid" else case["Yg"][:, :3]
    R, Val = host_bneu_assemble(kind, case["mesh"], IENb, gE, hg, Yg, case["rowPtr"], case["colPtr"], rho=1.06, bfs=0.2,
                                mvMsh=mvMsh, Do=Do, **TIME)
    assert np.abs(Rr).max() > 0 and (kind == "solid" or np.abs(Vr).max() > 0)
    assert np.array_equal(R, Rr) and np.array_equal(Val, Vr)


@pytest.mark.parametrize("elem,n", ELEMS)
def test_host_face_element_matches_golden(elem, n):
    g = golden("boundary_faces.npz")
    for kind, mvMsh in (("fluid", False), ("fluid", True), ("solid", False)):
        case, IENb, gE, hg, Do = _setup(elem, n, mvMsh)
        Yg = case["Yg"] if kind == "fluid" else case["Yg"][:, :3]
        R, Val = host_bneu_assemble(kind, case["mesh"], IENb, gE, hg, Yg, case["rowPtr"], case["colPtr"], rho=1.06, bfs=0.2,
                                    mvMsh=mvMsh, Do=Do, **TIME)
        tag = f"{elem}_{kind}_{int(mvMsh)}"
        assert rel_inf(R, g[f"R_{tag}"]) < 1e-14 and rel_inf(Val, g[f"Val_{tag}"]) < 1e-14


def _integ_inputs(elem, n):
    case, IENb, gE, hg, _ = _setup(elem, n)
    m = case["mesh"]
    rng = np.random.default_rng(12)
    Y = np.zeros((m.nNo, 7)); Y[:, :4] = case["Yg"]
    D = np.zeros((m.nNo, 7))
    D[:, 0:3] = 0.02 / n * rng.standard_normal((m.nNo, 3))
    D[:, 4:7] = 0.03 / n * rng.standard_normal((m.nNo, 3))
    return case, IENb, gE, Y, D


# (name, source rows l..u, geo, offset of the configuration rows): flux on the reference / old / new / moving configurations,
# pressure integral, area
INTEG = [("flux", 0, 2, 0, 0), ("flux_old", 0, 2, 1, 0), ("flux_new", 0, 2, 2, 0), ("flux_mv", 0, 2, 3, 4), ("pressure", 3, 3, 0, 0),
         ("area", None, None, 0, 0), ("area_mv", None, None, 3, 4)]


@pytest.mark.parametrize("elem,n", ELEMS)
@needs_ref
def test_host_face_integrals_match_reference_bitwise(elem, n):
    """all_fun::integ (all_fun.cpp:561,724,858): flux of the velocity, pressure integral and area of a face."""
    from oracle import ref
    case, IENb, gE, Y, D = _integ_inputs(elem, n)
    m = case["mesh"]
    ra = ref.RefAssembly(m.x, m.ien)
    for name, l, u, geo, goff in INTEG:
        want = ra.face_integ(IENb, gE, None if l is None else Y, 0 if l is None else l, u, geo=geo, D=D if geo else None)
        got = host_face_integ(m, IENb, gE, None if l is None else Y, 0 if l is None else l, u, geo=D if geo else None, goff=goff)
        assert got == want and want != 0.0, (name, got, want)
    assert abs(ra.face_integ(IENb, gE, None) - 1.0) < 1e-12          # Z0 of the unit block (boundary nodes are not jittered)
    ra.close()


@pytest.mark.gpu
@pytest.mark.parametrize("elem,n", ELEMS)
def test_gpu_face_integrals(elem, n):
    """b200_face_integ on the device-resident time-integrator arrays against the host run of the same arithmetic (which the
    CPU suite pins to the reference bit for bit) and the reference itself where it travelled: 1e-13 (the device contracts
    a*b + c into FMAs, the host build does not; the running sum has the reference's order on both)."""
    case, IENb, gE, Y, D = _integ_inputs(elem, n)
    m = case["mesh"]
    be = P.setup_backend(case)
    be.face_mesh_set(2, IENb, gE)
    be.pic_init(7, [dict(s=0, e=3, am=1.0, af=1.0, gam=1.0, beta=0.25), dict(s=4, e=6, am=1.0, af=1.0, gam=1.0, beta=0.25)], dFlag=True)
    be.pic_set("Yn", Y)
    from oracle import ref
    ra = ref.RefAssembly(m.x, m.ien) if ref.available() else None
    for name, l, u, geo, goff in INTEG:
        be.pic_set("Do", D if geo in (1, 3) else np.zeros_like(D))
        be.pic_set("Dn", D if geo == 2 else np.zeros_like(D))
        got = be.face_integ(2, None if l is None else "Yn", 0 if l is None else l, u, geo=geo)
        host = host_face_integ(m, IENb, gE, None if l is None else Y, 0 if l is None else l, u, geo=D if geo else None, goff=goff)
        assert abs(got - host) <= 1e-13 * abs(host), (name, got, host)
        if ra is not None:
            want = ra.face_integ(IENb, gE, None if l is None else Y, 0 if l is None else l, u, geo=geo, D=D if geo else None)
            assert abs(got - want) <= 1e-13 * abs(want), (name, got, want)
    be.close()


@pytest.mark.parametrize("elem,n", ELEMS)
@pytest.mark.parametrize("mvMsh", [False, True])
@needs_ref
def test_host_face_normals_match_reference_bitwise(elem, n, mvMsh):
    """fsi_ls_upd (eq_assem.cpp:316): val = int N_a n dGamma on the new-time-step / moving-mesh configuration."""
    from oracle import ref
    case, IENb, gE, Y, D = _integ_inputs(elem, n)
    m = case["mesh"]
    gN = np.unique(IENb)
    ra = ref.RefAssembly(m.x, m.ien)
    want = ra.fsi_ls_upd(IENb, gE, gN, D, mvMsh=mvMsh)
    ra.close()
    got = host_face_normals(m, IENb, gE, geo=D, goff=4 if mvMsh else 0)
    assert np.abs(want).max() > 0 and np.array_equal(got[gN], want)
    off = np.ones(m.nNo, bool); off[gN] = False
    assert np.abs(got[off]).max() == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("elem,n", ELEMS)
@pytest.mark.parametrize("geo", [2, 3])
def test_gpu_face_normal_update(elem, n, geo):
    """b200_face_normal_update: the coupled face's vector follows the moving configuration on the device."""
    case, IENb, gE, Y, D = _integ_inputs(elem, n)
    m = case["mesh"]
    gN = np.unique(IENb)
    be = P.setup_backend(case)                                   # faces 0..4 of the case are Dirichlet
    be.face_set(4, gN, 3, 1, np.zeros((len(gN), 3)), shared=False)      # turn slot 4 into the Neumann face Z0
    be.face_mesh_set(0, IENb, gE)
    be.pic_init(7, [dict(s=0, e=3, am=1.0, af=1.0, gam=1.0, beta=0.25), dict(s=4, e=6, am=1.0, af=1.0, gam=1.0, beta=0.25)], dFlag=True)
    be.pic_set("Dn" if geo == 2 else "Do", D)
    be.face_normal_update(0, 4, geo=geo)
    got = be.face_get_val(4, len(gN))
    host = host_face_normals(m, IENb, gE, geo=D, goff=4 if geo == 3 else 0)[gN]
    assert rel_inf(got, host) < 1e-13                     # FMA contraction on the device, none in the host build
    from oracle import ref
    if ref.available():
        ra = ref.RefAssembly(m.x, m.ien)
        assert rel_inf(got, ra.fsi_ls_upd(IENb, gE, gN, D, mvMsh=(geo == 3))) < 1e-13
        ra.close()
    be.close()


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("elem,n", ELEMS)
@pytest.mark.parametrize("kind,mvMsh", [("fluid", False), ("fluid", True), ("solid", False)])
def test_gpu_face_assembly_matches_golden(elem, n, kind, mvMsh):
    g = golden("boundary_faces.npz")
    case, IENb, gE, hg, Do = _setup(elem, n, mvMsh)
    be = P.setup_backend(case)
    be.face_mesh_set(5, IENb, gE)
    tDof = case["Yg"].shape[1]
    if kind == "fluid":
        be.state_set(tDof, case["Ag"], case["Yg"], case["Bf"])
        if mvMsh:
            be.disp_set(tDof, np.zeros_like(Do), Do)
        be.zero(4)
    else:
        be.zero(3)
    be.assemble_bneu(5, kind, hg, tDof=tDof, mvMsh=mvMsh, rho=1.06, bfs=0.2, **TIME)
    tag = f"{elem}_{kind}_{int(mvMsh)}"
    assert rel_inf(be.get_R(), g[f"R_{tag}"]) < 1e-12
    if kind == "fluid":
        assert rel_inf(be.get_Val(), g[f"Val_{tag}"]) < 1e-12
    else:
        assert np.abs(be.get_Val()).max() == 0.0
    be.close()


@pytest.mark.gpu
@pytest.mark.parametrize("elem,n", ELEMS)
def test_gpu_face_on_top_of_volume_assembly(elem, n):
    """Volume assembly, then the face: R / Val equal the reference's construct_fluid followed by b_assem_neu_bc."""
    g = golden("boundary_faces.npz")
    gv = golden("fluid_block.npz")
    case, IENb, gE, hg, Do = _setup(elem, n)
    be = P.setup_backend(case)
    be.face_mesh_set(0, IENb, gE)
    P.assemble(be, case)
    R0, V0 = be.get_R(), be.get_Val()
    be.assemble_bneu(0, "fluid", hg, tDof=4, rho=1.06, bfs=0.2, **TIME)
    R, Val = be.get_R(), be.get_Val()
    tag = f"{elem}_fluid_0"
    assert rel_inf(R - R0, g[f"R_{tag}"]) < 1e-10            # differences of assembled numbers: rounding of the larger terms
    assert rel_inf(R, R0 + g[f"R_{tag}"]) < 1e-12 and rel_inf(Val, V0 + g[f"Val_{tag}"]) < 1e-12
    if elem in ("hex", "tet10") and n == {"hex": 3, "tet10": 2}[elem]:
        key = "hex" if elem == "hex" else "tet10"
        assert rel_inf(R0, gv[f"R_{key}"]) < 1e-12
    be.close()


@pytest.mark.gpu
def test_gpu_unit_traction_integrates_to_area_normal():
    """Size-independent property on a large face (200 x 200 x 2 triangles): sum_a R(:,a) = -|face| n for h = 1."""
    m = M.block_mesh(20, "tet")
    case = P.fluid_block_case(20, elem="tet")
    on = np.abs(m.x[:, 0]) < 1e-12                             # X0, outward normal (-1, 0, 0)
    IENb, gE = M.face_elements(case["mesh"], on)
    be = P.setup_backend(case)
    be.face_mesh_set(1, IENb, gE)
    be.zero(3)
    be.assemble_bneu(1, "solid", np.ones(case["mesh"].nNo), tDof=3, **TIME)
    s = be.get_R().sum(axis=0)
    assert np.allclose(s, [1.0, 0.0, 0.0], atol=1e-13)
    be.close()


# ---------------------------------------------------------------------------------------------------------------------
# follower pressure load on a struct face (lBc.flwP): b_neu_folw_p + get_nnx / get_xi + b_struct_3d
# (eq_assem.cpp:186, nn.cpp:314-440, sv_struct.cpp:116)
# ---------------------------------------------------------------------------------------------------------------------
def _folw_setup(elem, n):
    case = P.block_case(n, elem=elem, kind="struct")
    m = case["mesh"]
    on = np.abs(m.x[:, 2] - 1.0) < 1e-12
    IENb, gE = M.face_elements(m, on)
    rng = np.random.default_rng(3)
    hg = np.where(on, 1.0e4 * (1.0 + 0.1 * rng.standard_normal(m.nNo)), 0.0)
    return case, IENb, gE, hg


@pytest.mark.parametrize("elem,n", ELEMS)
@needs_ref
def test_host_follower_pressure_matches_reference_bitwise(elem, n):
    from oracle import ref
    case, IENb, gE, hg = _folw_setup(elem, n)
    m, p = case["mesh"], case["props"]
    ra = ref.RefAssembly(m.x, m.ien)
    Rr, Vr = ra.bfolw(IENb, gE, hg, case["Dg"], dt=p["dt"], af=p["af"], beta=p["beta"])
    ra.close()
    R, Val = host_bfolw_assemble(m, IENb, gE, hg, case["Dg"], case["rowPtr"], case["colPtr"], dt=p["dt"], af=p["af"], beta=p["beta"])
    assert np.abs(Rr).max() > 0 and np.abs(Vr).max() > 0
    assert np.array_equal(R, Rr) and np.array_equal(Val, Vr)
    # the load follows the deformation: the tangent is not symmetric and vanishes on the diagonal entries of every block
    assert np.abs(Vr[:, [0, 4, 8]]).max() == 0.0


@pytest.mark.parametrize("elem,n", ELEMS)
def test_host_follower_pressure_matches_golden(elem, n):
    g = golden("boundary_faces.npz")
    case, IENb, gE, hg = _folw_setup(elem, n)
    m, p = case["mesh"], case["props"]
    R, Val = host_bfolw_assemble(m, IENb, gE, hg, case["Dg"], case["rowPtr"], case["colPtr"], dt=p["dt"], af=p["af"], beta=p["beta"])
    assert rel_inf(R, g[f"R_folw_{elem}"]) < 1e-14 and rel_inf(Val, g[f"Val_folw_{elem}"]) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("elem,n", ELEMS)
def test_gpu_follower_pressure_matches_golden(elem, n):
    g = golden("boundary_faces.npz")
    case, IENb, gE, hg = _folw_setup(elem, n)
    p = case["props"]
    be = P.setup_backend(case)
    be.face_mesh_set(3, IENb, gE)
    be.state_set(3, case["Ag"], case["Yg"], case["Bf"])
    be.disp_set(3, case["Dg"])
    be.zero(3)
    be.assemble_bfolw(3, hg, dt=p["dt"], af=p["af"], beta=p["beta"])
    assert rel_inf(be.get_R(), g[f"R_folw_{elem}"]) < 1e-12
    assert rel_inf(be.get_Val(), g[f"Val_folw_{elem}"]) < 1e-12
    # on top of the volume assembly: the sum, and a second call adds the same amount again (do_assem accumulates)
    P.assemble_solid(be, case)
    R0, V0 = be.get_R(), be.get_Val()
    be.assemble_bfolw(3, hg, dt=p["dt"], af=p["af"], beta=p["beta"])
    assert rel_inf(be.get_R(), R0 + g[f"R_folw_{elem}"]) < 1e-12 and rel_inf(be.get_Val(), V0 + g[f"Val_folw_{elem}"]) < 1e-12
    be.close()


def _folw_ustruct_setup(elem, n):
    case = P.ustruct_case(n, elem=elem)
    m = case["mesh"]
    on = np.abs(m.x[:, 2] - 1.0) < 1e-12
    IENb, gE = M.face_elements(m, on)
    rng = np.random.default_rng(3)
    hg = np.where(on, 1.0e4 * (1.0 + 0.1 * rng.standard_normal(m.nNo)), 0.0)
    return case, IENb, gE, hg


@pytest.mark.parametrize("elem,n", ELEMS)
@needs_ref
def test_host_ustruct_follower_pressure_matches_reference_bitwise(elem, n):
    """b_ustruct_3d + ustruct_do_assem (ustruct.cpp:132, 1579): R, the velocity block of Val and Kd."""
    from oracle import ref
    case, IENb, gE, hg = _folw_ustruct_setup(elem, n)
    m, p = case["mesh"], case["props"]
    kw = dict(dt=p["dt"], af=p["af"], ustruct=True, am=p["am"], gam=p["gam"])
    ra = ref.RefAssembly(m.x, m.ien)
    Rr, Vr, Kr = ra.bfolw(IENb, gE, hg, case["Dg"], **kw)
    ra.close()
    R, Val, Kd = host_bfolw_assemble(m, IENb, gE, hg, case["Dg"], case["rowPtr"], case["colPtr"], **kw)
    assert np.abs(Rr).max() > 0 and np.abs(Vr).max() > 0 and np.abs(Kr).max() > 0
    assert np.array_equal(R, Rr) and np.array_equal(Val, Vr) and np.array_equal(Kd, Kr)


@pytest.mark.gpu
@pytest.mark.parametrize("elem,n", ELEMS)
def test_gpu_ustruct_follower_pressure(elem, n):
    """On top of the ustruct volume assembly: R, Val and Kd grow by what the host run of the same arithmetic gives
    (pinned bit for bit to the reference on the CPU) -- and by the reference itself where it travelled."""
    case, IENb, gE, hg = _folw_ustruct_setup(elem, n)
    m, p = case["mesh"], case["props"]
    kw = dict(dt=p["dt"], af=p["af"], ustruct=True, am=p["am"], gam=p["gam"])
    Rh, Vh, Kh = host_bfolw_assemble(m, IENb, gE, hg, case["Dg"], case["rowPtr"], case["colPtr"], **kw)
    be = P.setup_backend(case)
    be.face_mesh_set(0, IENb, gE)
    P.assemble_ustruct(be, case)
    R0, V0, K0 = be.get_R(), be.get_Val(), be.get_Kd()
    be.assemble_bfolw(0, hg, tDof=4, **kw)
    assert rel_inf(be.get_R(), R0 + Rh) < 1e-12 and rel_inf(be.get_Val(), V0 + Vh) < 1e-12 and rel_inf(be.get_Kd(), K0 + Kh) < 1e-12
    assert rel_inf(be.get_Kd() - K0, Kh) < 1e-9              # the face's own share of Kd (difference of assembled numbers)
    from oracle import ref
    if ref.available():
        ra = ref.RefAssembly(m.x, m.ien)
        Rr, Vr, Kr = ra.bfolw(IENb, gE, hg, case["Dg"], **kw)
        ra.close()
        assert np.array_equal(Rh, Rr) and np.array_equal(Kh, Kr)
    be.close()
