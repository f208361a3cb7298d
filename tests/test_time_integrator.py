"""Generalised-alpha time integrator on the device (SURVEY.md par. 8(f) row 2): b200_picp / b200_pici / b200_picc against the
reference's own pic::picp / pici / picc (Code/Source/solver/pic.cpp:591,486,74; compiled into oracle/_ref), bit for bit, and a
whole Newton-converged time step that keeps the state on the device (no per-iteration upload of Ag/Yg, no download of the
solution) against the reference's time step at the north-star tolerance 1e-8."""
import numpy as np
import pytest

from conftest import needs_ref
from util import rel_l2

from svfsiplus_b200 import mesh as M
from svfsiplus_b200 import problem as P

FL = dict(s=0, e=3, am=0.0, af=0.0, gam=0.0, beta=0.0, phys="fluid", kind=0)


def _eq(phys, s, e, second_order=False, kind=0):
    if second_order:
        am, af, gam, beta = M.gen_alpha2(0.5)
    else:
        am, af, gam = M.gen_alpha(0.5)
        beta = 0.25 * (1.0 + am - af) ** 2
    return dict(s=s, e=e, am=am, af=af, gam=gam, beta=beta, phys=phys, kind=kind)


# (name, tDof, equations, dFlag, sstEq)
CONFIGS = [
    ("fluid", 4, [_eq("fluid", 0, 3)], False, False),
    ("struct", 3, [_eq("struct", 0, 2, second_order=True)], True, False),
    ("fluid+mesh", 7, [_eq("fluid", 0, 3), _eq("mesh", 4, 6, second_order=True)], True, False),
    ("ustruct", 4, [_eq("ustruct", 0, 3, kind=1)], True, True),
    ("ustruct+mesh", 7, [_eq("ustruct", 0, 3, kind=1), _eq("mesh", 4, 6, second_order=True, kind=0)], True, True),
]


def _state(nNo, tDof, seed=5):
    rng = np.random.default_rng(seed)
    st = {k: rng.standard_normal((nNo, tDof)) for k in ("Ao", "Yo", "Do", "An", "Yn", "Dn", "Ag", "Yg", "Dg")}
    st["Ad"] = rng.standard_normal((nNo, 3))
    return st


@needs_ref
def test_reference_pic_harness_formulas():
    """Pins the oracle harness: Bazilevs et al. 2007 eqs 86-90, 94-95 as pic.cpp writes them."""
    from oracle import ref
    st = _state(40, 7)
    eqs = CONFIGS[2][2]
    dt = 0.01
    o = ref.pic("p", st, eqs, dt=dt, dFlag=True)
    f, m = eqs
    assert np.array_equal(o["An"][:, :4], st["Ao"][:, :4] * ((f["gam"] - 1.0) / f["gam"]))
    assert np.array_equal(o["Yn"], st["Yo"])
    c = dt * dt * (0.5 * m["gam"] - m["beta"]) / (m["gam"] - 1.0)
    assert np.array_equal(o["Dn"][:, 4:], st["Do"][:, 4:] + o["Yn"][:, 4:] * dt + o["An"][:, 4:] * c)
    o = ref.pic("i", st, eqs, dt=dt)
    assert np.array_equal(o["Yg"][:, :4], st["Yo"][:, :4] * (1.0 - f["af"]) + st["Yn"][:, :4] * f["af"])
    R = np.random.default_rng(2).standard_normal((40, 4))
    o = ref.pic("c", st, eqs, dt=dt, cEq=0, R=R, Rd=np.zeros((40, 3)))
    assert np.array_equal(o["Yn"][:, :4], st["Yn"][:, :4] - R * (f["gam"] * dt))
    assert np.array_equal(o["Yn"][:, 4:], st["Yn"][:, 4:])


@pytest.mark.gpu
@pytest.mark.parametrize("name,tDof,eqs,dFlag,sstEq", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_pic_ops_bitwise(name, tDof, eqs, dFlag, sstEq):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not present on this box")
    case = P.pipe_case(3, 3, 4)
    be = P.setup_backend(case)
    nNo = be.nNo
    dt = 0.005
    st = _state(nNo, tDof)
    be.pic_init(tDof, eqs, dFlag=dFlag, sstEq=sstEq)
    for k in ("Ao", "Yo", "Do", "An", "Yn", "Dn", "Ad"):
        be.pic_set(k, st[k])
    # predictor
    be.picp(dt)
    r = ref.pic("p", st, eqs, dt=dt, dFlag=dFlag, sstEq=sstEq)
    for k in ("An", "Yn", "Dn", "Ad"):
        assert np.array_equal(be.pic_get(k), r[k]), (name, "picp", k)
    st = r
    # Dirichlet values written into the predicted state (set_bc_dir)
    idx = np.array([0, tDof + 1, 5 * tDof + 2], np.int32)
    val = np.array([1.5, -2.5, 3.5])
    be.pic_scatter("Yn", idx, val)
    st["Yn"].reshape(-1)[idx] = val
    # initiator
    be.pici()
    r = ref.pic("i", st, eqs, dt=dt, dFlag=dFlag, sstEq=sstEq)
    for k in ("Ag", "Yg", "Dg"):
        got, want = be.pic_get(k), r[k]
        for q in eqs:                                  # rows outside every equation are not written by either side
            assert np.array_equal(got[:, q["s"]:q["e"] + 1], want[:, q["s"]:q["e"] + 1]), (name, "pici", k)
    st = r
    # corrector, every equation in turn, first and later Newton iterations
    rng = np.random.default_rng(11)
    for first in (True, False):
        for iEq, q in enumerate(eqs):
            dof = q["e"] - q["s"] + 1
            R = rng.standard_normal((nNo, dof))
            be.zero(dof)
            be.set_R(R)
            be.picc(iEq, dt, first_itr=first)
            if q["kind"] == 1 and first:                # ustruct_r (ustruct.cpp:1753-1764)
                amg = (q["gam"] - q["am"]) / (q["gam"] - 1.0)
                Rd = amg * st["Ad"] - st["Yg"][:, q["s"]:q["s"] + 3]
            else:
                Rd = np.zeros((nNo, 3))
            r = ref.pic("c", st, eqs, dt=dt, cEq=iEq, dFlag=dFlag, sstEq=sstEq, R=R, Rd=Rd)
            for k in ("An", "Yn", "Dn", "Ad"):
                assert np.array_equal(be.pic_get(k), r[k]), (name, "picc", iEq, first, k)
            st = r
    # end of the time step
    be.pic_advance()
    for o, n in (("Ao", "An"), ("Yo", "Yn"), ("Do", "Dn")):
        assert np.array_equal(be.pic_get(o), st[n])
    be.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ls", ["NS", "GMRES"])
def test_time_step_device_resident(ls):
    """picp -> [pici -> ls_alloc -> construct_fluid -> fsils_solve -> picc] x n with the state resident on the device,
    against the same loop made of the reference's own functions: velocity / pressure / acceleration within 1e-8."""
    from oracle import ref, refcase
    if not ref.available():
        pytest.skip("oracle/_ref not present on this box")
    case = P.pipe_case(8, 8, 16, coupled=False)
    p = case["props"]
    dt, n_newton = p["dt"], 6
    eqs = [dict(s=0, e=3, am=p["am"], af=p["af"], gam=p["gam"], beta=0.0, phys="fluid", kind=0)]
    nNo = case["mesh"].nNo
    zeros = np.zeros((nNo, 4))

    # device
    be = P.setup_backend(case)
    be.pic_init(4, eqs)
    be.pic_set("Ao", case["Ag"]); be.pic_set("Yo", case["Yg"]); be.pic_set("Do", zeros)
    be.picp(dt)
    be.state_set(4, None, None, case["Bf"])             # body force once; Ag / Yg never leave the device
    norms_g = []
    for it in range(n_newton):
        be.pici()
        X, info = P.newton_linear_step(be, case, ls=ls, upload=False, fetch=False)
        norms_g.append(info["RI"]["iNorm"])
        be.picc(0, dt, first_itr=(it == 0))
    An_g, Yn_g = be.pic_get("An"), be.pic_get("Yn")
    be.close()

    # reference
    st = dict(Ao=case["Ag"], Yo=case["Yg"], Do=zeros, An=zeros, Yn=zeros, Dn=zeros, Ad=np.zeros((nNo, 3)), Ag=zeros, Yg=zeros, Dg=zeros)
    st = ref.pic("p", st, eqs, dt=dt)
    norms_r = []
    for it in range(n_newton):
        st = ref.pic("i", st, eqs, dt=dt)
        c = dict(case); c["Ag"] = st["Ag"]; c["Yg"] = st["Yg"]
        R, Val, X, o = refcase.reference_step(c, ls)
        norms_r.append(o["iNorm"])
        st = ref.pic("c", st, eqs, dt=dt, R=X, Rd=np.zeros((nNo, 3)))
    assert norms_r[-1] < 1e-9 * norms_r[0] and norms_g[-1] < 1e-9 * norms_g[0]
    assert rel_l2(Yn_g[:, :3], st["Yn"][:, :3]) < 1e-8
    assert rel_l2(Yn_g[:, 3], st["Yn"][:, 3]) < 1e-8
    assert rel_l2(An_g, st["An"]) < 1e-8
