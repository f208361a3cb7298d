"""Golden fixtures for the features added late in round 1 (after the round's GPU budget was spent), generated from the
COMPILED REFERENCE (oracle/_ref) like make_golden.py:
    python tests/golden/make_golden_late.py
Writes tests/golden/late_additions.npz; the GPU tests that read it live in tests/test_zz_late_additions.py (sorted last so a
failure there cannot mask the suites that were green on the B200)."""
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
    out = {}
    # Holzapfel-Ogden with modified anisotropy (stIso_HO_ma): struct (get_pk2cc<3>) and ustruct (get_pk2cc_dev)
    for elem in ("tet", "hex"):
        c = P.block_case(3, elem=elem, kind="struct", iso="HO_ma", vol="ST91")
        R, Val, *_ = refcase.reference_assemble_solid(c)
        out[f"R_{elem}_struct_HO_ma"] = R
        out[f"Val_{elem}_struct_HO_ma"] = Val
        c = P.ustruct_case(3, elem=elem, iso="HO_ma")
        R, Val, Kd, _ = refcase.reference_assemble_ustruct(c)
        out[f"R_{elem}_ustruct_HO_ma"] = R
        out[f"Val_{elem}_ustruct_HO_ma"] = Val
        out[f"Kd_{elem}_ustruct_HO_ma"] = Kd
    # solid viscosity (dmn.solid_visc: Newtonian and pseudo-potential models) in struct_3d
    for elem, n in (("tet", 3), ("hex", 3), ("tet10", 2)):
        for visc in ("newt", "pot"):
            c = P.block_case(n, elem=elem, kind="struct", iso="nHook", vol="ST91", visc=visc, visc_mu=5.0e4)
            R, Val, *_ = refcase.reference_assemble_solid(c)
            out[f"R_{elem}_struct_visc_{visc}"] = R
            out[f"Val_{elem}_struct_visc_{visc}"] = Val
    # prestress: S += S0 from the nodal com_mod.pS0, and the pstEq accumulations pSn / pSa (with and without viscosity)
    for elem, n in (("tet", 3), ("hex", 3), ("tet10", 2)):
        for visc in (None, "pot"):
            c = P.block_case(n, elem=elem, kind="struct", iso="nHook", vol="ST91", visc=visc, visc_mu=5.0e4, prestress=True)
            R, Val, *_ = refcase.reference_assemble_solid(c)
            tag = f"{elem}_struct_pst_{visc}"
            out[f"R_{tag}"], out[f"Val_{tag}"], out[f"pSn_{tag}"], out[f"pSa_{tag}"] = R, Val, c["_ref_pSn"], c["_ref_pSa"]
    # FSI with a prestressed, viscous wall (construct_fsi reads com_mod.pS0; dmn.solid_visc of the struct domain)
    for tag, c in (("tet", P.fsi_case(4, 4, 4)), ("hex", P.fsi_block_case(3, elem="hex")), ("tet10", P.fsi_block_case(2, elem="tet10"))):
        rng = np.random.default_rng(77)
        c["solid"] = dict(c["solid"], visc="pot", visc_mu=200.0)
        c["pS0"] = 1.0e4 * rng.standard_normal((c["mesh"].nNo, 6))
        R, Val, _ = refcase.reference_assemble_fsi(c)
        out[f"R_{tag}_fsi_wall"], out[f"Val_{tag}_fsi_wall"] = R, Val
    # ... and in ustruct_3d_m (Siso + Svis, Kvis_u in Ku, af Kvis_v)
    for elem, n in (("tet", 3), ("hex", 3), ("tet10", 2)):
        for visc in ("newt", "pot"):
            c = P.ustruct_case(n, elem=elem, visc=visc, visc_mu=5.0e4)
            R, Val, Kd, _ = refcase.reference_assemble_ustruct(c)
            out[f"R_{elem}_ustruct_visc_{visc}"] = R
            out[f"Val_{elem}_ustruct_visc_{visc}"] = Val
            out[f"Kd_{elem}_ustruct_visc_{visc}"] = Kd
    np.savez_compressed(os.path.join(HERE, "late_additions.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
