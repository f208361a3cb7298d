"""Displacement-based and mixed solid equations on quadratic tetrahedra (TET10, 15 Gauss points, curved edges): the K11
kernels k_assemble_solid / k_assemble_ustruct instantiated for 10 nodes, through the C ABI, against the compiled reference
(construct_dsolid, construct_l_elas, construct_mesh, construct_usolid on the same mesh) at 1e-12."""
import numpy as np
import pytest

from conftest import needs_ref
from util import golden, rel_inf

from svfsiplus_b200 import problem as P

TOL_ASM = 1e-12
SOLID = [("struct", "nHook", "ST91"), ("struct", "HO", "ST91"), ("struct", "mStVK", None), ("lelas", None, None), ("mesh", None, None)]


def _solid_case(kind, iso, vol):
    return P.block_case(2, elem="tet10", kind=kind, iso=iso or "nHook", vol=vol)


@needs_ref
def test_oracle_reproduces_tet10_fixtures():
    from oracle import refcase
    g = golden("block_tet10.npz")
    for kind, iso, vol in SOLID:
        R, Val, _, _, _, _ = refcase.reference_assemble_solid(_solid_case(kind, iso, vol))
        assert np.array_equal(R, g[f"R_{kind}_{iso}_{vol}"]) and np.array_equal(Val, g[f"Val_{kind}_{iso}_{vol}"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind,iso,vol", SOLID)
def test_tet10_solid_assembly_matches_golden(kind, iso, vol):
    g = golden("block_tet10.npz")
    case = _solid_case(kind, iso, vol)
    be = P.setup_backend(case)
    P.assemble_solid(be, case)
    assert rel_inf(be.get_R(), g[f"R_{kind}_{iso}_{vol}"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{kind}_{iso}_{vol}"]) < TOL_ASM
    be.close()


@pytest.mark.gpu
@pytest.mark.parametrize("iso", ["nHook", "HO"])
def test_tet10_ustruct_assembly_matches_golden(iso):
    """construct_usolid + ustruct_do_assem (R, Val, Kd) and ustruct_r on equal-order TET10."""
    g = golden("block_tet10.npz")
    case = P.ustruct_case(2, elem="tet10", iso=iso)
    be = P.setup_backend(case)
    P.assemble_ustruct(be, case)
    assert rel_inf(be.get_R(), g[f"uR_{iso}"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"uVal_{iso}"]) < TOL_ASM
    assert rel_inf(be.get_Kd(), g[f"uKd_{iso}"]) < TOL_ASM
    P.assemble_ustruct(be, case, upload=False, with_r=True)
    assert rel_inf(be.get_R(), g[f"uRr_{iso}"]) < TOL_ASM
    be.close()


# ---------------------------------------------------------------------------------------------------------------------
# fibre-reinforcement / active stress (stM.Tf: get_fib_stress, mat_models.cpp:126; Tfa along the fibre, Tfa*eta_s along the
# sheet direction; mat_models_carray.h:222-225, 386, 984, 1015; mat_models.cpp:682-684, 704, 914, 929)
# ---------------------------------------------------------------------------------------------------------------------
ACTIVE = [("struct", "hex", "HO"), ("struct", "tet", "nHook"), ("struct", "tet10", "HO"), ("ustruct", "tet", "HO"), ("ustruct", "hex", "nHook")]


def _active_case(eq, elem, iso):
    n = 2 if elem == "tet10" else 3
    if eq == "struct":
        case = P.block_case(n, elem=elem, kind="struct", iso="HO", vol="ST91")      # brings fibre directions
        if iso == "nHook":
            E = 240.56596e6
            case["props"].update(iso="nHook", C10=0.25 * E / 1.5, Kpen=4.0e9)
            case["props"].pop("ho")
        case["props"].update(Tfa=3.0e4, eta_s=0.4)
    else:
        case = P.ustruct_case(n, elem=elem, iso="HO")
        if iso == "nHook":
            case["props"].pop("ho"); case["props"]["iso"] = "nHook"
        case["props"].update(Tfa=2.0e4, eta_s=0.3)
    return case


@needs_ref
def test_oracle_reproduces_active_stress_fixtures():
    from oracle import refcase
    g = golden("active_stress.npz")
    for eq, elem, iso in ACTIVE:
        case = _active_case(eq, elem, iso)
        if eq == "struct":
            R, Val, _, _, _, _ = refcase.reference_assemble_solid(case)
            zero = dict(case); zero["props"] = dict(case["props"], Tfa=0.0)
            R0, _, _, _, _, _ = refcase.reference_assemble_solid(zero)
        else:
            R, Val, Kd, _ = refcase.reference_assemble_ustruct(case)
            zero = dict(case); zero["props"] = dict(case["props"], Tfa=0.0)
            R0, _, _, _ = refcase.reference_assemble_ustruct(zero)
        assert np.array_equal(R, g[f"R_{eq}_{elem}_{iso}"]) and np.array_equal(Val, g[f"Val_{eq}_{elem}_{iso}"])
        assert rel_inf(R0, R) > 1e-6, (eq, elem, iso)           # the fibre stress matters


@pytest.mark.gpu
@pytest.mark.parametrize("eq,elem,iso", ACTIVE)
def test_active_fibre_stress_matches_golden(eq, elem, iso):
    g = golden("active_stress.npz")
    case = _active_case(eq, elem, iso)
    be = P.setup_backend(case)
    if eq == "struct":
        P.assemble_solid(be, case)
    else:
        P.assemble_ustruct(be, case)
        assert rel_inf(be.get_Kd(), g[f"Kd_{eq}_{elem}_{iso}"]) < TOL_ASM
    assert rel_inf(be.get_R(), g[f"R_{eq}_{elem}_{iso}"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{eq}_{elem}_{iso}"]) < TOL_ASM
    be.close()


# ---------------------------------------------------------------------------------------------------------------------
# Mooney-Rivlin, Holzapfel-Gasser-Ogden and Guccione laws (mat_models_carray.h:438-540, 544-688, 692-903; the laws themselves are pinned on the CPU
# in tests/test_solid_laws.py)
# ---------------------------------------------------------------------------------------------------------------------
NEW_LAWS = [(iso, elem) for iso in ("MR", "HGO", "Gucci") for elem in ("tet", "hex", "tet10")]


def _law_case(iso, elem):
    return P.block_case(2 if elem == "tet10" else 3, elem=elem, kind="struct", iso=iso, vol="ST91")


@needs_ref
def test_oracle_reproduces_new_law_fixtures():
    from oracle import refcase
    g = golden("active_stress.npz")
    for iso, elem in NEW_LAWS:
        R, Val, _, _, _, _ = refcase.reference_assemble_solid(_law_case(iso, elem))
        assert np.array_equal(R, g[f"R_{iso}_{elem}"]) and np.array_equal(Val, g[f"Val_{iso}_{elem}"])


@pytest.mark.gpu
@pytest.mark.parametrize("iso,elem", NEW_LAWS)
def test_new_law_assembly_matches_golden(iso, elem):
    g = golden("active_stress.npz")
    case = _law_case(iso, elem)
    be = P.setup_backend(case)
    P.assemble_solid(be, case)
    assert rel_inf(be.get_R(), g[f"R_{iso}_{elem}"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"Val_{iso}_{elem}"]) < TOL_ASM
    be.close()


# the same laws in the mixed (ustruct) formulation: get_pk2cc_dev (mat_models.cpp:630) = the law without volumetric terms
USTRUCT_LAWS = [("MR", "tet"), ("HGO", "hex"), ("Gucci", "tet"), ("Gucci", "tet10")]


@needs_ref
def test_oracle_reproduces_ustruct_law_fixtures():
    from oracle import refcase
    g = golden("active_stress.npz")
    for iso, elem in USTRUCT_LAWS:
        R, Val, Kd, _ = refcase.reference_assemble_ustruct(P.ustruct_case(2 if elem == "tet10" else 3, elem=elem, iso=iso))
        assert np.array_equal(R, g[f"uR_{iso}_{elem}"]) and np.array_equal(Val, g[f"uVal_{iso}_{elem}"])


@pytest.mark.gpu
@pytest.mark.parametrize("iso,elem", USTRUCT_LAWS)
def test_ustruct_law_assembly_matches_golden(iso, elem):
    g = golden("active_stress.npz")
    case = P.ustruct_case(2 if elem == "tet10" else 3, elem=elem, iso=iso)
    be = P.setup_backend(case)
    P.assemble_ustruct(be, case)
    assert rel_inf(be.get_R(), g[f"uR_{iso}_{elem}"]) < TOL_ASM
    assert rel_inf(be.get_Val(), g[f"uVal_{iso}_{elem}"]) < TOL_ASM
    assert rel_inf(be.get_Kd(), g[f"uKd_{iso}_{elem}"]) < TOL_ASM
    be.close()
