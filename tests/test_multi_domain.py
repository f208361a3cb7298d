"""Equations with several domains of one physics (eq.nDmn > 1): every element takes the properties of its own domain
(all_fun::domain, Code/Source/solver/all_fun.cpp:149; construct_fluid fluid.cpp:531, construct_dsolid sv_struct.cpp:261).
Device: one divergence-free launch per domain over that domain's element list (b200_mesh_domains +
b200_assemble_fluid_dmn / b200_assemble_struct_dmn), against the compiled reference at 1e-12.

Reference defect not copied: construct_dsolid keeps a COPY of com_mod.cDmn (sv_struct.cpp:229) where construct_fluid takes a
reference (fluid.cpp:491), so struct_3d_carray reads one domain's properties for every element.  The struct fixture is
therefore the reference assembled one domain at a time (the other domains skipped by its own phys test) and added."""
import numpy as np
import pytest

from conftest import needs_ref
from util import golden, rel_inf

from svfsiplus_b200 import backend as B
from svfsiplus_b200 import problem as P


def struct_setup(elem="hex", n=3):
    case = P.block_case(n, elem=elem, kind="struct", iso="HO", vol="ST91")          # brings fibre directions
    m = case["mesh"]
    cen = m.x[m.ien].mean(axis=1)
    elem_dmn = np.minimum((cen[:, 0] * 3.0).astype(np.int32), 2)                    # three slabs in x
    base = {k: case["props"][k] for k in ("dt", "am", "af", "gam", "beta")}
    E = 240.56596e6
    mu = 0.5 * E / 1.5
    props = [dict(base, rho=1000.0, iso="nHook", vol="ST91", C10=0.5 * mu, Kpen=4.0e9, f=(0.0, 0.0, -9.81)),
             dict(base, rho=1200.0, dmp=3.0, iso="mStVK", vol=None, C10=E / 1.2, C01=0.5 * E / 1.3),
             dict(base, rho=1060.0, iso="HO", vol="M94", Kpen=1.0e6, ho=case["props"]["ho"])]
    return case, elem_dmn, props


def fluid_setup(elem="tet", n=3):
    case = P.fluid_block_case(n, elem=elem)
    m = case["mesh"]
    cen = m.x[m.ien].mean(axis=1)
    elem_dmn = (cen[:, 2] > 0.5).astype(np.int32)
    base = {k: case["props"][k] for k in ("dt", "am", "af", "gam")}
    props = [dict(base, rho=1.06, mu=0.04, f=(0.1, 0.0, 0.0)),
             dict(base, rho=1.2, mu=0.04, viscType=1, mu_o=0.6, lam=8.2, a=1.23, n=0.64, Kinv=0.5)]
    return case, elem_dmn, props


@needs_ref
def test_oracle_reproduces_multi_domain_fixtures():
    from oracle import ref
    g = golden("multi_domain.npz")
    case, ed, props = struct_setup()
    ra = ref.RefAssembly(case["mesh"].x, case["mesh"].ien)
    ra.set_fibers(case["fN"])
    R, Val = ra.struct_domains(ed, props, case["Ag"], case["Yg"], case["Dg"], case["Bf"])
    assert np.array_equal(R, g["R_struct"]) and np.array_equal(Val, g["Val_struct"])
    # the domains matter: a single-domain assembly with the first domain's law differs
    R1, V1, _ = ra.solid("struct", case["Ag"], case["Yg"], case["Dg"], case["Bf"], **props[0])
    assert rel_inf(V1, Val) > 1e-3
    # ... and the three domains are additive pieces of three single-domain assemblies
    on0 = np.isin(np.arange(case["mesh"].nNo), np.unique(case["mesh"].ien[ed == 0]))
    only0 = ~np.isin(np.arange(case["mesh"].nNo), np.unique(case["mesh"].ien[ed != 0]))
    assert only0.any() and np.array_equal(R[only0], R1[only0])
    ra.close()
    case, ed, props = fluid_setup()
    ra = ref.RefAssembly(case["mesh"].x, case["mesh"].ien)
    R, Val = ra.fluid_domains(ed, props, case["Ag"], case["Yg"], case["Bf"])
    assert np.array_equal(R, g["R_fluid"]) and np.array_equal(Val, g["Val_fluid"])
    ra.close()


@pytest.mark.gpu
def test_struct_three_domains_matches_golden():
    g = golden("multi_domain.npz")
    case, ed, props = struct_setup()
    be = P.setup_backend(case)
    be.state_set(3, case["Ag"], case["Yg"], case["Bf"])
    be.disp_set(3, case["Dg"])
    be.mesh_fibers(case["fN"])
    be.mesh_domains(3, ed)
    be.zero(3)
    be.assemble_struct_dmn([B.struct_props(tDof=3, **p) for p in props])
    assert rel_inf(be.get_R(), g["R_struct"]) < 1e-12
    assert rel_inf(be.get_Val(), g["Val_struct"]) < 1e-12
    be.close()


@pytest.mark.gpu
def test_fluid_two_domains_matches_golden():
    g = golden("multi_domain.npz")
    case, ed, props = fluid_setup()
    be = P.setup_backend(case)
    be.state_set(4, case["Ag"], case["Yg"], case["Bf"])
    be.mesh_domains(2, ed)
    be.zero(4)
    be.assemble_fluid_dmn([B.fluid_props(tDof=4, **p) for p in props])
    assert rel_inf(be.get_R(), g["R_fluid"]) < 1e-11          # pow() of the Carreau-Yasuda domain: a few ulp vs libm
    assert rel_inf(be.get_Val(), g["Val_fluid"]) < 1e-11
    be.close()
