"""Workloads of the hot path: the pipe_RCR_3d Navier-Stokes case on synthetic refined pipes.

Everything the reference reads from tests/cases/fluid/pipe_RCR_3d/solver.xml is restated here as
data (density 1.06, constant viscosity 0.04, dt 0.005, rho_inf 0.5, <LS type="NS"> tol 1e-3, 15 outer
iterations, GM 1e-3/10, CG 1e-3/300, absTol 1e-17, Krylov dimension 250, preconditioner fsils; inlet
and wall Dirichlet, outlet RCR-coupled Neumann).  The mesh is generated (svfsiplus_b200/mesh.py); the
outlet resistance enters the linear solve as res = gam*dt*r (Code/Source/solver/main.cpp:571-578).
"""
from __future__ import annotations

import numpy as np

from . import backend as B
from . import mesh as M

# <LS> blocks: (ls_type, RI, GM, CG) with tolerances (relTol, absTol, mItr, sD)
LS_SETTINGS = {
    # tests/cases/fluid/pipe_RCR_3d/solver.xml:75-87
    "NS": (B.LS_NS, (1e-3, 1e-17, 15, 250), (1e-3, 1e-17, 10, 250), (1e-3, 1e-17, 300, 0)),
    # plain GMRES variant of SURVEY.md §8d (the metric names GMRES): same tolerance, restarts of 250
    "GMRES": (B.LS_GMRES, (1e-3, 1e-17, 4, 250), None, None),
    "CG": (B.LS_CG, (1e-3, 1e-12, 50, 0), None, None),
    "BICGS": (B.LS_BICGS, (1e-8, 1e-14, 200, 0), None, None),
    # tests/cases/struct/block_compression/solver.xml: <LS type="BICG"> tol 1e-12, 600 iterations
    "BICGS_STRUCT": (B.LS_BICGS, (1e-12, 1e-10, 600, 0), None, None),
    "GMRES_STRUCT": (B.LS_GMRES, (1e-9, 1e-10, 10, 100), None, None),
    "GMRES_STRUCT_LOOSE": (B.LS_GMRES, (1e-4, 1e-10, 10, 100), None, None),
    # tests/cases/fsi/pipe_3d/solver.xml: FSI equation <LS type="GMRES"> tol 1e-12, 100 iterations, Krylov dim 50
    "GMRES_FSI": (B.LS_GMRES, (1e-12, 1e-10, 100, 50), None, None),
    # tests/cases/ustruct/block_compression/P1P1_VMS/solver.xml: <LS type="GMRES"> tol 1e-12 (Krylov dim capped here)
    "GMRES_USTRUCT": (B.LS_GMRES, (1e-8, 1e-10, 10, 200), None, None),
    "GMRES_USTRUCT_LOOSE": (B.LS_GMRES, (1e-3, 1e-10, 10, 200), None, None),
    "CG_MESH": (B.LS_CG, (1e-10, 1e-14, 400, 0), None, None),
}


def pipe_case(nx, ny, nz, *, radius=1.0, length=10.0, jitter=0.1, coupled=True, visc=None, pattern=None):
    """pattern: optional callable (nNo, ien) -> (rowPtr, colPtr), e.g. the device-side lhsa (Backend.pattern); default is
    the host construction M.csr_pattern (~30 s at 10 M tets)."""
    m = M.pipe_mesh(nx, ny, nz, radius=radius, length=length, jitter=jitter)
    rowPtr, colPtr = pattern(m.nNo, m.ien) if pattern else M.csr_pattern(m.ien, m.nNo)
    am, af, gam = M.gen_alpha(0.5)
    Ag, Yg, Bf = M.pipe_state(m, radius=radius, length=length)
    dt = 0.005
    props = dict(dt=dt, am=am, af=af, gam=gam, rho=1.06, mu=0.04)
    if visc:
        props.update(visc)
    out_nodes = m.faces["outlet"]["nodes"]
    faces = [
        dict(name="lumen_inlet", nodes=m.faces["inlet"]["nodes"], dof=3, bGrp=B.BC_DIR,
             val=np.zeros((len(m.faces["inlet"]["nodes"]), 3))),
        dict(name="lumen_wall", nodes=m.faces["wall"]["nodes"], dof=3, bGrp=B.BC_DIR,
             val=np.zeros((len(m.faces["wall"]["nodes"]), 3))),
        dict(name="lumen_outlet", nodes=out_nodes, dof=3, bGrp=B.BC_NEU,
             val=M.face_normal_integral(m.x, m.faces["outlet"]["tris"], out_nodes)),
    ]
    r_out = 121.0 + 1212.0                          # Rp + Rd of the RCR block
    res = np.array([0.0, 0.0, gam * dt * r_out if coupled else 0.0])
    incL = np.array([1, 1, 1], np.int32)
    return dict(mesh=m, rowPtr=rowPtr, colPtr=colPtr, Ag=Ag, Yg=Yg, Bf=Bf, props=props, faces=faces,
                res=res, incL=incL, name=f"pipe_{nx}x{ny}x{nz}")


def fluid_block_case(n, elem="hex", *, visc=None, Kinv=0.0, mvMsh=False):
    """Navier-Stokes on the unit block meshed with HEX8 ("hex"), TET10 ("tet10", curved edges) or TET4 ("tet"):
    the element types of SURVEY.md par. 8 rows A4/A5 beyond the linear tet.  Fluid parameters of pipe_RCR_3d; a
    smooth swirling velocity + noise, body force and (optionally) Darcy permeability so every term of fluid_3d_m/c is
    exercised.  Walls X0, X1, Y0, Y1 and the inflow Z0 are Dirichlet, Z1 is a traction-free outflow."""
    m = M.block_mesh(n, elem=elem)
    rowPtr, colPtr = M.csr_pattern(m.ien, m.nNo)
    am, af, gam = M.gen_alpha(0.5)
    x = m.x
    nN = m.nNo
    tDof = 7 if mvMsh else 4
    rng = np.random.default_rng(3030)
    Yg = np.zeros((nN, tDof)); Ag = np.zeros((nN, tDof))
    Yg[:, 0] = 5.0 * np.sin(np.pi * x[:, 1]) * np.cos(np.pi * x[:, 2])
    Yg[:, 1] = -5.0 * np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 2])
    Yg[:, 2] = 20.0 * x[:, 0] * (1.0 - x[:, 0]) * x[:, 1] * (1.0 - x[:, 1])
    Yg[:, :3] += 0.2 * rng.standard_normal((nN, 3))
    Yg[:, 3] = 100.0 * (1.0 - x[:, 2]) + rng.standard_normal(nN)
    Ag[:, :4] = 10.0 * rng.standard_normal((nN, 4))
    if mvMsh:
        Yg[:, 4:7] = 0.5 * rng.standard_normal((nN, 3))
    Bf = 0.5 * rng.standard_normal((nN, 3))
    props = dict(dt=0.005, am=am, af=af, gam=gam, rho=1.06, mu=0.04, f=(0.3, -0.2, 0.1), Kinv=Kinv, mvMsh=mvMsh)
    if visc:
        props.update(visc)
    faces = []
    for nm in ("X0", "X1", "Y0", "Y1", "Z0"):
        nodes = m.faces[nm]["nodes"]
        faces.append(dict(name=nm, nodes=nodes, dof=3, bGrp=B.BC_DIR, val=np.zeros((len(nodes), 3))))
    return dict(mesh=m, rowPtr=rowPtr, colPtr=colPtr, Ag=Ag, Yg=Yg, Bf=Bf, props=props, faces=faces,
                res=np.zeros(len(faces)), incL=np.ones(len(faces), np.int32), name=f"fluid_block_{elem}_{n}")


def setup_backend(case, device=0, be=None) -> B.Backend:
    """Single-rank set-up: what initialize() + fsi_ls_ini + add_eq_linear_algebra do once."""
    m = case["mesh"]
    if be is None:
        be = B.Backend(device)
    be.lhs_create(m.nNo, case["rowPtr"], case["colPtr"], nFaces=len(case["faces"]))
    for i, f in enumerate(case["faces"]):
        be.face_set(i, f["nodes"], f["dof"], f["bGrp"], f["val"], shared=False)
    be.mesh_set(m.ien, m.x)
    return be


def assemble(be: B.Backend, case, upload=True):
    """ls_alloc + global_eq_assem for the fluid equation."""
    if upload:
        be.state_set(case["Ag"].shape[1], case["Ag"], case["Yg"], case["Bf"])
    be.zero(4)
    p = case["props"]
    be.assemble_fluid(B.fluid_props(tDof=case["Ag"].shape[1], **p))
    if case.get("nranks", 1) > 1:
        be.commu_R()                  # all_fun::commu(R), main.cpp:513


def newton_linear_step(be: B.Backend, case, ls="NS", want_system=False, upload=True, fetch=True, out=None,
                       prec=B.PREC_FSILS):
    """One Newton iteration's hot path: assemble R/Val, then solve.  Returns (X, info[, R, Val])."""
    assemble(be, case, upload=upload)
    R = Val = None
    if want_system:
        R, Val = be.get_R(), be.get_Val()
    ls_type, RI, GM, CG = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    X, info = be.solve(ls_type, prec, RI, GM, CG, case["incL"], case["res"], out=out, fetch=fetch)
    if want_system:
        return X, info, R, Val
    return X, info


# ---------------------------------------------------------------------------------------------------
# solid block (tests/cases/struct/block_compression/solver.xml: neo-Hookean, E 240.56596e6, nu 0.5, ST91
# penalty 4e9, density 1000, dt 1e-4, rho_inf 0.5; X0/Y0/Z0 Dirichlet in one direction each)
# ---------------------------------------------------------------------------------------------------
def block_case(n, elem="hex", kind="struct", iso="nHook", vol="ST91", jitter=0.1, visc=None, visc_mu=0.0, prestress=False, pattern=None):
    """pattern: optional callable (nNo, ien) -> (rowPtr, colPtr), e.g. the device-side lhsa (Backend.pattern)."""
    m = M.block_mesh(n, elem=elem, jitter=jitter)
    rowPtr, colPtr = pattern(m.nNo, m.ien) if pattern else M.csr_pattern(m.ien, m.nNo)
    am, af, gam, beta = M.gen_alpha2(0.5)
    Ag, Yg, Dg, Bf = M.block_state(m)
    E, nu = 240.56596e6, 0.5
    mu = 0.5 * E / (1.0 + nu)
    dt = 1e-4
    if kind == "struct":
        if iso == "nHook":
            C10, C01 = 0.5 * mu, 0.0
        elif iso in ("HO", "HO_ma", "HGO", "Gucci"):      # Holzapfel-Ogden myocardium (parameters of tests/cases/struct/LV_* style, cgs)
            C10, C01 = 0.0, 0.0
        elif iso == "MR":                        # Mooney-Rivlin: C10 + C01 = mu / 2
            C10, C01 = 0.3 * mu, 0.2 * mu
        elif iso == "StVK":                      # C10 = lambda, C01 = mu (nu 0.3 keeps lambda finite)
            C10, C01 = E * 0.3 / (1.3 * 0.4), 0.5 * E / 1.3
        else:                                    # mStVK: C10 = kappa, C01 = mu
            C10, C01 = E / (3.0 * 0.4), 0.5 * E / 1.3
        props = dict(dt=dt, am=am, af=af, gam=gam, beta=beta, rho=1000.0, dmp=0.0, f=(0.0, 0.0, 0.0), iso=iso, vol=vol,
                     C10=C10, C01=C01, Kpen=4.0e9 if vol else 0.0)
        if iso in ("HO", "HO_ma"):
            props["ho"] = dict(a=590.0, b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12, afs=2160.0, bfs=11.436, khs=100.0)
            props["Kpen"] = 1.0e6
        if visc:                                 # solid viscosity (tests/cases/struct/tensile_adventitia_*_viscosity: mu 50)
            props.update(visc=visc, visc_mu=visc_mu)
        if iso == "Gucci":                       # Guccione myocardium: C10 and the three exponents
            props.update(C10=880.0, Kpen=1.0e6, ho=dict(bff=8.0, bss=6.0, bfs=12.0))
        if iso == "HGO":                         # arterial-wall style parameters: two dispersed fibre families
            props.update(C10=0.5 * mu, kap=0.226, ho=dict(aff=9.96e5, bff=524.6, ass=9.96e5, bss=524.6))
    elif kind == "lelas":
        props = dict(dt=dt, am=am, af=af, gam=gam, beta=beta, rho=1000.0, elM=E, nu=0.3, f=(0.0, 0.0, -9.81))
    else:
        # ALE mesh-motion equation as the FSI case configures it: unknowns in rows 4..6 of a tDof = 7 state
        # (fluid/struct velocity-pressure first), unit modulus, reference configuration x + Do
        props = dict(dt=dt, am=am, af=af, gam=gam, beta=beta, rho=0.0, elM=1.0, nu=0.3, f=(0.0, 0.0, 0.0), s=4)
        A7, Y7, D7, _ = M.block_state(m, tDof=7, s=4)
        rng = np.random.default_rng(77)
        Do = np.zeros_like(D7)
        Do[:, 4:7] = 0.5 * D7[:, 4:7] + 0.002 * rng.standard_normal((m.nNo, 3))
        Ag, Yg, Dg = A7, Y7, D7
    faces = []
    for ax, nm in enumerate(("X0", "Y0", "Z0")):
        nodes = m.faces[nm]["nodes"]
        val = np.ones((len(nodes), 3))
        val[:, ax] = 0.0
        faces.append(dict(name=nm, nodes=nodes, dof=3, bGrp=B.BC_DIR, val=val))
    case = dict(mesh=m, rowPtr=rowPtr, colPtr=colPtr, Ag=Ag, Yg=Yg, Dg=Dg, Bf=Bf, props=props, faces=faces, kind=kind,
                res=np.zeros(len(faces)), incL=np.ones(len(faces), np.int32), name=f"block_{elem}_{n}_{kind}")
    if kind == "struct" and iso in ("HO", "HO_ma", "HGO", "Gucci"):
        # fibre / sheet directions rotating through the block (unit, orthogonal), one pair per element
        cen = m.x[m.ien].mean(axis=1)
        th = 0.5 * np.pi * cen[:, 2] + 0.3 * cen[:, 0]
        fN = np.zeros((m.nEl, 6))
        fN[:, 0], fN[:, 1] = np.cos(th), np.sin(th)
        fN[:, 3], fN[:, 4] = -np.sin(th), np.cos(th)
        case["fN"] = fN
    if kind == "mesh":
        case["Do"] = Do
    if prestress:
        # a prestress field of the size of the elastic stresses of this state (symmetric tensor per node, rows 00 11 22 01 12 20)
        # and the pstEq accumulations switched on (com_mod.pS0 / pstEq, sv_struct.cpp:232-235)
        rng = np.random.default_rng(3131)
        case["pS0"] = 2.0e6 * rng.standard_normal((m.nNo, 6))
        case["pstEq"] = True
    return case


def assemble_solid(be: B.Backend, case, upload=True):
    """ls_alloc + global_eq_assem for a struct / lElas equation (dof 3)."""
    tDof = case["Ag"].shape[1]
    if upload:
        be.state_set(tDof, case["Ag"], case["Yg"], case["Bf"])
        be.disp_set(tDof, case["Dg"], case.get("Do"))
    if upload and case.get("fN") is not None:
        be.mesh_fibers(case["fN"])
    if upload and (case.get("pS0") is not None or case.get("pstEq")):
        be.prestress_set(case.get("pS0"), case.get("pstEq", False))          # com_mod.pS0 / pstEq
    be.zero(3)
    if case["kind"] == "struct":
        be.assemble_struct(B.struct_props(tDof=tDof, **case["props"]))
    else:
        be.assemble_lelas(B.lelas_props(tDof=tDof, mesh_mode=(case["kind"] == "mesh"), **case["props"]))
    if case.get("nranks", 1) > 1:
        be.commu_R()


def solid_linear_step(be: B.Backend, case, ls="BICGS_STRUCT", want_system=False, prec=B.PREC_FSILS):
    assemble_solid(be, case)
    R = Val = None
    if want_system:
        R, Val = be.get_R(), be.get_Val()
    ls_type, RI, GM, CG = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    X, info = be.solve(ls_type, prec, RI, GM, CG, case["incL"], case["res"])
    return (X, info, R, Val) if want_system else (X, info)


# ---------------------------------------------------------------------------------------------------
# FSI pipe (tests/cases/fsi/pipe_3d/solver.xml): lumen = fluid domain, outer shell of the same structured
# pipe = struct domain (neo-Hookean wall), one dof-4 matrix, tDof = 7 (FSI unknowns 0..3, mesh 4..6)
# ---------------------------------------------------------------------------------------------------
def fsi_case(nx, ny, nz, *, wall_from=0.8, radius=1.0, length=10.0, pattern=None):
    m = M.pipe_mesh(nx, ny, nz, radius=radius, length=length)
    rowPtr, colPtr = pattern(m.nNo, m.ien) if pattern else M.csr_pattern(m.ien, m.nNo)
    am, af, gam = M.gen_alpha(0.5)                 # FSI is a first-order equation (initialize.cpp:424-433)
    beta = 0.25 * (1.0 + am - af) ** 2
    dt = 1e-4
    cen = m.x[m.ien].mean(axis=1)
    r = np.sqrt(cen[:, 0] ** 2 + cen[:, 1] ** 2)
    elem_dmn = (r > wall_from * radius).astype(np.int32)       # 0 fluid (lumen), 1 struct (wall)
    n = m.nNo
    A4, Y4, Bf = M.pipe_state(m, radius=radius, length=length)
    rng = np.random.default_rng(99)
    Ag = np.zeros((n, 7)); Yg = np.zeros((n, 7)); Dg = np.zeros((n, 7))
    Ag[:, :4] = A4; Yg[:, :4] = Y4
    # displacement noise relative to the LOCAL spacing (the disc map's cells near the square's corners are far smaller than
    # 2R/nx on fine meshes; a global amplitude inverts them).  Small meshes keep the global amplitude their goldens were made with.
    h = 2.0 * radius / nx
    big = nx * ny * nz > 2000                                   # (the small parity cases keep the state their goldens were made with)
    if big:
        h = np.minimum(M.node_hmin(m.x, m.ien), h)[:, None]
    Dg[:, 0:3] = 0.02 * h * rng.standard_normal((n, 3))         # solid displacement (used in the wall)
    Dg[:, 4:7] = 0.05 * h * rng.standard_normal((n, 3))         # mesh displacement (used in the lumen)
    if big:
        # flat cells along the map's diagonals must not fold in either displaced configuration: halve the displacement of the nodes
        # of any element that loses more than half of its volume until none does (what pipe_mesh does for its jitter)
        v0 = M.tet_volumes(m.x, m.ien)
        for sl in (slice(0, 3), slice(4, 7)):
            for _ in range(20):
                v = M.tet_volumes(m.x + Dg[:, sl], m.ien)
                bad = (v * v0 <= 0.0) | (np.abs(v) < 0.5 * np.abs(v0))
                if not bad.any():
                    break
                Dg[np.unique(m.ien[bad].reshape(-1)), sl] *= 0.5
    Yg[:, 4:7] = 0.5 * rng.standard_normal((n, 3))              # mesh velocity
    Ag[:, 4:7] = rng.standard_normal((n, 3))
    Bf = 0.1 * rng.standard_normal((n, 3))
    E, nu = 1.0e7, 0.3
    mu = 0.5 * E / (1.0 + nu)
    fluid = dict(rho=1.0, mu=0.04)
    solid = dict(rho=1.0, dmp=0.0, iso="nHook", vol="ST91", C10=0.5 * mu, C01=0.0, Kpen=E / (3.0 * (1.0 - 2.0 * nu)))
    faces = [
        dict(name="inlet", nodes=m.faces["inlet"]["nodes"], dof=3, bGrp=B.BC_DIR, val=np.zeros((len(m.faces["inlet"]["nodes"]), 3))),
        dict(name="outlet", nodes=m.faces["outlet"]["nodes"], dof=3, bGrp=B.BC_DIR, val=np.zeros((len(m.faces["outlet"]["nodes"]), 3))),
    ]
    return dict(mesh=m, rowPtr=rowPtr, colPtr=colPtr, Ag=Ag, Yg=Yg, Dg=Dg, Bf=Bf, elem_dmn=elem_dmn, fluid=fluid, solid=solid,
                time=dict(dt=dt, am=am, af=af, gam=gam, beta=beta), faces=faces, res=np.zeros(2), incL=np.ones(2, np.int32),
                kind="fsi", name=f"fsi_{nx}x{ny}x{nz}")


def fsi_block_case(n, elem="hex"):
    """construct_fsi on the unit block (HEX8 or TET4): x < 0.5 is the fluid domain on the ALE configuration, the rest
    a neo-Hookean struct domain; same unknown layout as fsi_case (tDof 7)."""
    m = M.block_mesh(n, elem=elem)
    rowPtr, colPtr = M.csr_pattern(m.ien, m.nNo)
    am, af, gam = M.gen_alpha(0.5)
    beta = 0.25 * (1.0 + am - af) ** 2
    cen = m.x[m.ien].mean(axis=1)
    elem_dmn = (cen[:, 0] > 0.5).astype(np.int32)
    base = fluid_block_case(n, elem=elem, mvMsh=True)
    nN = m.nNo
    h = 1.0 / n
    rng = np.random.default_rng(98)
    Dg = np.zeros((nN, 7))
    Dg[:, 0:3] = 0.02 * h * rng.standard_normal((nN, 3))
    Dg[:, 4:7] = 0.05 * h * rng.standard_normal((nN, 3))
    E, nu = 1.0e7, 0.3
    mu = 0.5 * E / (1.0 + nu)
    fluid = dict(rho=1.0, mu=0.04)
    solid = dict(rho=1.0, dmp=0.0, iso="nHook", vol="ST91", C10=0.5 * mu, C01=0.0, Kpen=E / (3.0 * (1.0 - 2.0 * nu)))
    faces = [dict(name=nm, nodes=m.faces[nm]["nodes"], dof=3, bGrp=B.BC_DIR, val=np.zeros((len(m.faces[nm]["nodes"]), 3)))
             for nm in ("Z0", "Z1")]
    return dict(mesh=m, rowPtr=rowPtr, colPtr=colPtr, Ag=base["Ag"], Yg=base["Yg"], Dg=Dg, Bf=base["Bf"], elem_dmn=elem_dmn,
                fluid=fluid, solid=solid, time=dict(dt=1e-4, am=am, af=af, gam=gam, beta=beta), faces=faces, res=np.zeros(2),
                incL=np.ones(2, np.int32), kind="fsi", name=f"fsi_block_{elem}_{n}")


def assemble_fsi(be: B.Backend, case, upload=True):
    t = case["time"]
    if upload:
        be.state_set(7, case["Ag"], case["Yg"], case["Bf"])
        be.disp_set(7, case["Dg"])
        be.mesh_domains(2, case["elem_dmn"])
        if case.get("pS0") is not None:
            be.prestress_set(case["pS0"], False)              # the wall's prestress: read, never accumulated (fsi.cpp:147-148)
    be.zero(4)
    fp = B.fluid_props(tDof=7, mvMsh=True, dt=t["dt"], am=t["am"], af=t["af"], gam=t["gam"], **case["fluid"])
    sp = B.struct_props(tDof=7, s=0, dt=t["dt"], am=t["am"], af=t["af"], gam=t["gam"], beta=t["beta"], **case["solid"])
    be.assemble_fsi([0, 1], [fp, None], [None, sp])


def fsi_linear_step(be: B.Backend, case, ls="GMRES_FSI", want_system=False):
    assemble_fsi(be, case)
    R = Val = None
    if want_system:
        R, Val = be.get_R(), be.get_Val()
    ls_type, RI, GM, CG = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    X, info = be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case["incL"], case["res"])
    return (X, info, R, Val) if want_system else (X, info)


# ---------------------------------------------------------------------------------------------------
# ustruct block (tests/cases/ustruct/block_compression/P1P1_VMS/solver.xml: neo-Hookean, E 240.56596e6,
# nu 0.4999999, ST91, density 1e-3, stabilisation coefficients 1e-3, first-order generalised-alpha)
# ---------------------------------------------------------------------------------------------------
def ustruct_case(n, elem="tet", vol="ST91", iso="nHook", visc=None, visc_mu=0.0, pattern=None):
    m = M.block_mesh(n, elem=elem)
    rowPtr, colPtr = pattern(m.nNo, m.ien) if pattern else M.csr_pattern(m.ien, m.nNo)
    am, af, gam = M.gen_alpha(0.5)
    E, nu = 240.56596e6, 0.4999999
    mu = 0.5 * E / (1.0 + nu)
    kap = E / (3.0 * (1.0 - 2.0 * nu))
    dt = 1e-3
    A3, Y3, D3, Bf = M.block_state(m)
    rng = np.random.default_rng(515)
    nN = m.nNo
    Ag = np.zeros((nN, 4)); Yg = np.zeros((nN, 4)); Dg = np.zeros((nN, 4))
    Ag[:, :3] = A3; Yg[:, :3] = 0.1 * Y3; Dg[:, :3] = 0.2 * D3
    Yg[:, 3] = 1.0e3 * rng.standard_normal(nN)        # pressure
    Ag[:, 3] = 1.0e4 * rng.standard_normal(nN)        # pressure rate
    Ad = rng.standard_normal((nN, 3))                 # displacement-equation acceleration (com_mod.Ad)
    # stabilisation coefficients 0.1 instead of the XML's 1e-3: on this random (non-equilibrium) state the Jacobian with
    # 1e-3 needs ~1600 GMRES iterations in the reference itself, with 0.1 about 200
    props = dict(dt=dt, am=am, af=af, gam=gam, rho=1e-3, elM=E, nu=nu, ctM=0.1, ctC=0.1, vol=vol, C10=0.5 * mu, Kpen=kap,
                 f=(0.0, 0.0, -9.81))
    faces = []
    for ax, nm in enumerate(("X0", "Y0", "Z0")):
        nodes = m.faces[nm]["nodes"]
        val = np.ones((len(nodes), 3))
        val[:, ax] = 0.0
        faces.append(dict(name=nm, nodes=nodes, dof=3, bGrp=B.BC_DIR, val=val))
    case = dict(mesh=m, rowPtr=rowPtr, colPtr=colPtr, Ag=Ag, Yg=Yg, Dg=Dg, Bf=Bf, Ad=Ad, props=props, faces=faces, kind="ustruct",
                res=np.zeros(len(faces)), incL=np.ones(len(faces), np.int32), name=f"ustruct_{elem}_{n}")
    if visc:                                     # solid viscosity (tests/cases/ustruct/tensile_adventitia_*_viscosity)
        props.update(visc=visc, visc_mu=visc_mu)
    if iso != "nHook":
        props["iso"] = iso
        if iso in ("HO", "HO_ma"):
            props["ho"] = dict(a=590.0, b=8.023, aff=184720.0, bff=16.026, ass=24810.0, bss=11.12, afs=2160.0, bfs=11.436, khs=100.0)
        elif iso == "MR":
            props.update(C10=0.3 * mu, C01=0.2 * mu)
        elif iso == "HGO":
            props.update(kap=0.226, ho=dict(aff=9.96e5, bff=524.6, ass=9.96e5, bss=524.6))
        elif iso == "Gucci":
            props.update(C10=880.0, ho=dict(bff=8.0, bss=6.0, bfs=12.0))
        if iso != "MR":                          # fibre / sheet directions rotating through the block
            cen = m.x[m.ien].mean(axis=1)
            th = 0.5 * np.pi * cen[:, 2] + 0.3 * cen[:, 0]
            fN = np.zeros((m.nEl, 6))
            fN[:, 0], fN[:, 1] = np.cos(th), np.sin(th)
            fN[:, 3], fN[:, 4] = -np.sin(th), np.cos(th)
            case["fN"] = fN
    return case


def assemble_ustruct(be: B.Backend, case, upload=True, with_r=False):
    if upload:
        be.state_set(4, case["Ag"], case["Yg"], case["Bf"])
        be.disp_set(4, case["Dg"])
        if case.get("fN") is not None:
            be.mesh_fibers(case["fN"])
    be.zero(4)
    p = case["props"]
    be.assemble_ustruct(B.ustruct_props(tDof=4, **p))
    if with_r:                                         # main.cpp:526, first Newton iteration
        amg = (p["gam"] - p["am"]) / (p["gam"] - 1.0)
        be.ustruct_r(amg, 1.0 / p["am"], 0, case["Ad"])


def ustruct_linear_step(be: B.Backend, case, ls="GMRES_USTRUCT", with_r=True):
    assemble_ustruct(be, case, with_r=with_r)
    R, Val, Kd = be.get_R(), be.get_Val(), be.get_Kd()
    ls_type, RI, GM, CG = LS_SETTINGS[ls] if isinstance(ls, str) else ls
    X, info = be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case["incL"], case["res"])
    return X, info, R, Val, Kd
