"""Drop-in boundary test (`-m gpu`): the plug-in class svfsiplus_b200/host/B200LinearAlgebra.cpp, compiled
against the UNMODIFIED reference headers, is driven through the reference's own objects (ComMod, eqType,
FSILS_lhsType built by fsils_lhs_create / fsils_bc_create) and the reference's own call sequence
ls_alloc -> global assembly -> ls_solve (Code/Source/solver/ls.cpp:51-82), and compared with the
reference's FsilsLinearAlgebra path on the same inputs."""
import numpy as np
import pytest

from util import rel_l2

from svfsiplus_b200 import backend as B
from svfsiplus_b200 import problem as P

pytestmark = pytest.mark.gpu


def _need():
    from oracle import ref
    if not (ref.available() and ref.dropin_available()):
        pytest.skip("oracle/_ref (libsvref.so + libsvdropin.so) not present on this box")


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("ls", ["NS", "GMRES"])
def test_plugin_class_matches_fsils_backend(mode, ls):
    _need()
    from oracle import ref, refcase
    case = P.pipe_case(8, 8, 16)
    lsv = refcase._ls_vector(P.LS_SETTINGS[ls])
    X, info = ref.dropin_fluid_step(case, lsv, mode)
    assert int(info["device_assembly"]) == mode
    Rr, Vr, Xr, oref = refcase.reference_step(case, ls)
    assert rel_l2(X, Xr) < 1e-5          # production tolerance 1e-3 solve, see test_gpu_parity.TOL_SOL_LOOSE
    assert abs(int(info["itr"]) - int(oref["itr"])) <= 1
    assert bool(info["suc"]) == bool(oref["suc"])


def test_plugin_class_tight_solve():
    _need()
    from oracle import ref, refcase
    from svfsiplus_b200 import backend as B
    case = P.pipe_case(8, 8, 16)
    ls = (B.LS_GMRES, (1e-11, 1e-30, 10, 300), None, None)
    X, info = ref.dropin_fluid_step(case, refcase._ls_vector(ls), 1)
    Rr, Vr, Xr, oref = refcase.reference_step(case, ls)
    assert rel_l2(X, Xr) < 1e-6        # cond x 1e-11, see test_gpu_parity.test_tight_tolerance_solution


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("elem,kind,ls", [("hex", "struct", "BICGS_STRUCT"), ("tet", "struct", "BICGS_STRUCT"),
                                          ("tet", "lelas", "GMRES_STRUCT_LOOSE")])
def test_plugin_class_solid_equations(mode, elem, kind, ls):
    """struct / lElas through the reference's own ComMod with eq.linear_algebra = B200LinearAlgebra: host assembly +
    device solve (mode 0) and device assembly through assemble_mesh (mode 1), against the FsilsLinearAlgebra path."""
    _need()
    from oracle import ref, refcase
    case = P.block_case(6, elem=elem, kind=kind)
    X, info = ref.dropin_solid_step(case, refcase._ls_vector(P.LS_SETTINGS[ls]), mode)
    assert int(info["device_assembly"]) == mode
    Rr, Vr, Xr, oref = refcase.reference_solid_step(case, ls)
    assert bool(info["suc"]) == bool(oref["suc"])
    assert rel_l2(X, Xr) < (1e-8 if ls.startswith("BICGS") else 1e-2)
    assert abs(int(info["itr"]) - int(oref["itr"])) <= max(1, 0.02 * oref["itr"])


@pytest.mark.parametrize("mode", [0, 1])
def test_plugin_class_neumann_face(mode):
    """A Neumann face (traction + backflow stabilisation) after the volume assembly: with device assembly it goes through
    B200LinearAlgebra::assemble_face (no per-element assemble() call), otherwise through the reference's b_assem_neu_bc;
    either way the step matches the reference's construct_fluid + b_assem_neu_bc + fsils_solve."""
    _need()
    from oracle import ref, refcase
    from svfsiplus_b200 import mesh as M
    case = P.fluid_block_case(6, elem="tet")
    m = case["mesh"]
    on = np.abs(m.x[:, 2] - 1.0) < 1e-12                    # the traction-free face Z1
    IENb, gE = M.face_elements(m, on)
    hg = np.where(on, 40.0 + 5.0 * np.sin(3.0 * m.x[:, 0]), 0.0)
    # reverse the axial velocity on part of the face so that the backflow term is active
    case["Yg"][:, 2] *= np.where(m.x[:, 0] < 0.5, -1.0, 1.0)
    case["props"]["f"] = (0.0, 0.0, 0.0)                     # the harness configures the fluid without a body force
    ls = (B.LS_GMRES, (1e-3, 1e-14, 10, 150), None, None)
    X, info = ref.dropin_fluid_face_step(case, refcase._ls_vector(ls), mode, IENb, gE, hg)
    assert int(info["device_assembly"]) == mode and int(info["face_on_device"]) == mode
    R, Val, rowPtr, colPtr, _ = refcase.reference_assemble(case)
    ra = ref.RefAssembly(m.x, m.ien)
    p = case["props"]
    Rf, Vf = ra.bneu("fluid", IENb, gE, hg, case["Yg"], dt=p["dt"], af=p["af"], gam=p["gam"], rho=p["rho"], bfs=0.2)
    ra.close()
    assert np.abs(Vf).max() > 0.0
    Xr, oref = refcase.reference_solve(case, R + Rf, Val + Vf, ls)
    assert abs(int(info["itr"]) - int(oref["itr"])) <= 1
    assert rel_l2(X, Xr) < 1e-8
