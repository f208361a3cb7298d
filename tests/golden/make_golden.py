"""Generates the golden fixtures under tests/golden/ from the COMPILED REFERENCE (oracle/_ref).

Run here (where /root/reference exists and `make -C oracle ref` has been run):
    python tests/golden/make_golden.py
The fixtures pin the oracle and the CUDA path to outputs of the reference's own code
(construct_fluid + fsils_solve) on seeded synthetic inputs; they travel with the repo so the GPU box
needs neither /root/reference nor the oracle build.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refcase  # noqa: E402
from svfsiplus_b200 import problem as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # full system + all four solvers on a tiny pipe
    case = P.pipe_case(4, 4, 6)
    R, Val, rowPtr, colPtr, _ = refcase.reference_assemble(case)
    out = dict(R=R, Val=Val, rowPtr=rowPtr, colPtr=colPtr)
    for ls in ("NS", "GMRES", "CG", "BICGS"):
        X, o = refcase.reference_solve(case, R, Val, P.LS_SETTINGS[ls])
        out[f"X_{ls}"] = X
        out[f"info_{ls}"] = np.array([o["suc"], o["itr"], o["iNorm"], o["fNorm"], o["GM_itr"], o["CG_itr"]])
    # non-Newtonian viscosity models + moving-mesh convective velocity (assembly only)
    for tag, visc in (("cy", dict(viscType=1, mu=0.04, mu_o=0.6, lam=8.2, a=1.23, n=0.64)),
                      ("cass", dict(viscType=2, mu=0.3, mu_o=0.4, lam=0.5))):
        c2 = P.pipe_case(4, 4, 6, visc=visc)
        R2, V2, _, _, _ = refcase.reference_assemble(c2)
        out[f"R_{tag}"] = R2
        out[f"Val_{tag}"] = V2
    np.savez_compressed(os.path.join(HERE, "pipe_4_4_6.npz"), **out)

    case = P.pipe_case(6, 6, 12)
    R, Val, X, o = refcase.reference_step(case, "NS")
    np.savez_compressed(os.path.join(HERE, "pipe_6_6_12_ns.npz"), X=X,
                        info=np.array([o["suc"], o["itr"], o["iNorm"], o["fNorm"], o["GM_itr"], o["CG_itr"]]))
    # solid equations on a 3^3 block: struct (three laws), lElas, mesh; TET4 and HEX8
    out = {}
    for elem in ("tet", "hex"):
        for kind, iso, vol in (("struct", "nHook", "ST91"), ("struct", "nHook", "M94"), ("struct", "StVK", None),
                               ("struct", "mStVK", None), ("struct", "HO", "ST91"), ("lelas", None, None), ("mesh", None, None)):
            c = P.block_case(3, elem=elem, kind=kind, iso=iso or "nHook", vol=vol)
            R, Val, _, _, _, tabs = refcase.reference_assemble_solid(c)
            tag = f"{elem}_{kind}_{iso}_{vol}"
            out[f"R_{tag}"] = R
            out[f"Val_{tag}"] = Val
        out[f"w_{elem}"], out[f"N_{elem}"], out[f"Nx_{elem}"] = tabs
    np.savez_compressed(os.path.join(HERE, "block_3_solid.npz"), **out)
    # FSI equation (construct_fsi): fluid lumen + struct wall on one dof-4 matrix
    c = P.fsi_case(4, 4, 4)
    R, Val, _ = refcase.reference_assemble_fsi(c)
    np.savez_compressed(os.path.join(HERE, "fsi_4_4_4.npz"), R=R, Val=Val, elem_dmn=c["elem_dmn"])
    # ustruct equation (construct_usolid + ustruct_r): R, Val, Kd
    out = {}
    for elem in ("tet", "hex"):
        for vol in ("ST91", "M94", "Quad"):
            c = P.ustruct_case(3, elem=elem, vol=vol)
            R, Val, Kd, _ = refcase.reference_assemble_ustruct(c, with_r=False)
            Rr, _, _, _ = refcase.reference_assemble_ustruct(c, with_r=True)
            out[f"R_{elem}_{vol}"] = R; out[f"Val_{elem}_{vol}"] = Val; out[f"Kd_{elem}_{vol}"] = Kd; out[f"Rr_{elem}_{vol}"] = Rr
        c = P.ustruct_case(3, elem=elem, iso="HO")
        R, Val, Kd, _ = refcase.reference_assemble_ustruct(c, with_r=False)
        out[f"R_{elem}_HO"] = R; out[f"Val_{elem}_HO"] = Val; out[f"Kd_{elem}_HO"] = Kd
    np.savez_compressed(os.path.join(HERE, "ustruct_3.npz"), **out)
    # generic fluid element (HEX8, TET10 with curved edges): assembly for every branch + one tight GMRES step
    import test_fluid_elements as T
    from svfsiplus_b200 import backend as B
    out = {}
    for tag, elem, n, kw in T.GOLDEN_CASES:
        R, Val, _, _, _ = refcase.reference_assemble(P.fluid_block_case(n, elem=elem, **kw))
        out[f"R_{tag}"] = R; out[f"Val_{tag}"] = Val
    for tag, elem, n in (("hex", "hex", 6), ("tet10", "tet10", 3)):
        R, Val, X, o = refcase.reference_step(P.fluid_block_case(n, elem=elem), T.LS_STEP)
        out[f"X_step_{tag}"] = X
        out[f"info_step_{tag}"] = np.array([o["suc"], o["itr"], o["iNorm"], o["fNorm"]])
    for elem, n in (("hex", 4), ("tet10", 2)):
        R, Val, _ = refcase.reference_assemble_fsi(P.fsi_block_case(n, elem=elem))
        out[f"R_fsi_{elem}"] = R; out[f"Val_fsi_{elem}"] = Val
    np.savez_compressed(os.path.join(HERE, "fluid_block.npz"), **out)
    # solid equations on curved TET10 (15 Gauss points)
    import test_tet10_solids as T10
    out = {}
    for kind, iso, vol in T10.SOLID:
        R, Val, _, _, _, _ = refcase.reference_assemble_solid(T10._solid_case(kind, iso, vol))
        out[f"R_{kind}_{iso}_{vol}"] = R; out[f"Val_{kind}_{iso}_{vol}"] = Val
    for iso in ("nHook", "HO"):
        c = P.ustruct_case(2, elem="tet10", iso=iso)
        R, Val, Kd, _ = refcase.reference_assemble_ustruct(c, with_r=False)
        Rr, _, _, _ = refcase.reference_assemble_ustruct(c, with_r=True)
        out[f"uR_{iso}"] = R; out[f"uVal_{iso}"] = Val; out[f"uKd_{iso}"] = Kd; out[f"uRr_{iso}"] = Rr
    np.savez_compressed(os.path.join(HERE, "block_tet10.npz"), **out)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
