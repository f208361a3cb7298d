#!/usr/bin/env python
"""Profiling drivers (run on the GPU box, normally under ncu):

  python tools/prof.py tour  [--dims nx ny nz] [--reps R]   stand-alone bench of every kernel class
  python tools/prof.py step  [--dims nx ny nz] [--ls NS]    one resident Newton-iteration hot path
  python tools/prof.py asm   [--dims nx ny nz] [--reps R]   assembly only
  python tools/prof.py solid [--reps R]                     K11 solid element kernels on production-size blocks
  python tools/prof.py fluidgen [--reps R]                  K10 generic fluid element (HEX8 100^3, TET10 6 x 32^3)

`tour` prints one JSON line per kernel class (ms, algorithmic GB/s, fraction of the measured HBM peak);
under `ncu --set full -k regex:k_` it gives one capture per kernel on production-size data.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svfsiplus_b200 import backend as B  # noqa: E402
from svfsiplus_b200 import problem as P  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["tour", "tiled", "step", "asm", "solid", "fluidgen"])
    ap.add_argument("--dims", type=int, nargs=3, default=[96, 96, 181])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--ls", default="NS")
    a = ap.parse_args()
    if a.mode == "solid":
        # K11 element kernels on production-size blocks: struct HEX8 100^3 (1.0 M hex), struct / ustruct TET4 60^3 x 6
        for name, case, asm in (("struct_hex8_100", P.block_case(100, elem="hex", kind="struct"), P.assemble_solid),
                                ("struct_tet4_60", P.block_case(60, elem="tet", kind="struct"), P.assemble_solid),
                                ("ustruct_tet4_60", P.ustruct_case(60, elem="tet"), P.assemble_ustruct)):
            be = P.setup_backend(case)
            asm(be, case)
            be.timer_start()
            for _ in range(a.reps):
                asm(be, case, upload=False)
            ms = be.timer_stop() / a.reps
            nEl = case["mesh"].nEl
            print(json.dumps(dict(kernel=name, nEl=nEl, ms=ms, ns_per_elem=1e6 * ms / nEl, nnz=be.nnz)), flush=True)
            be.close()
        return
    if a.mode == "fluidgen":
        for name, case in (("fluid_hex8_100", P.fluid_block_case(100, elem="hex")), ("fluid_tet10_32", P.fluid_block_case(32, elem="tet10"))):
            be = P.setup_backend(case)
            P.assemble(be, case)
            if a.reps > 1:
                P.assemble(be, case, upload=False)
            be.timer_start()
            for _ in range(a.reps):
                P.assemble(be, case, upload=False)
            ms = be.timer_stop() / a.reps
            nEl = case["mesh"].nEl
            print(json.dumps(dict(kernel=name, nEl=nEl, nNo=be.nNo, nnz=be.nnz, ms=ms, ns_per_elem=1e6 * ms / nEl,
                                  val_write_GBps=be.nnz * 128.0 / 1e6 / ms)), flush=True)
            be.close()
        return
    t0 = time.time()
    case = P.pipe_case(*a.dims)
    be = P.setup_backend(case)
    print(f"# setup {time.time()-t0:.1f} s: nNo {be.nNo} nnz {be.nnz} nEl {case['mesh'].nEl}", flush=True)
    pk = peak()
    if a.mode == "asm":
        be.state_set(case["Ag"].shape[1], case["Ag"], case["Yg"], case["Bf"])
        props = B.fluid_props(tDof=case["Ag"].shape[1], **case["props"])
        for _ in range(2):
            be.zero(4); be.assemble_fluid(props)
        be.timer_start()
        for _ in range(a.reps):
            be.zero(4); be.assemble_fluid(props)
        ms = be.timer_stop() / a.reps
        by = be.nnz * 128.0 + be.nNo * (32.0 + 24 + 64 + 24) + case["mesh"].nEl * 16.0
        print(json.dumps(dict(kernel="assembly(+zero)", ms=ms, GBps=by / 1e6 / ms, frac=by / 1e6 / ms / pk,
                              ns_per_tet=1e6 * ms / case["mesh"].nEl)))
        return
    P.assemble(be, case)
    if a.mode == "tiled":
        # A/B of the TMA-staged row-tile kernels (k = 2) against the per-lane kernels on the same data
        plan = [("spmv_vv3", 0), ("spmv_vv3", 1), ("spmv_vv3", 2), ("spmv_vv3", 3), ("spmv_vv3", 4), ("spmv_vv3", 5), ("spmv_vv3", 6), ("spmv_sv", 0), ("spmv_sv", 3), ("spmv_sv", 2), ("spmv_vs", 1), ("spmv_vs", 2),
                ("spmv_ss", 0), ("spmv_ss", 2)]
        for name, k in plan:
            ms, by = be.op_bench(name, k=k, reps=a.reps)
            print(json.dumps(dict(kernel=name, k=k, ms=ms, MB=by / 1e6, GBps=by / 1e6 / ms, frac=by / 1e6 / ms / pk)), flush=True)
        be.close()
        return
    if a.mode == "tour":
        plan = [("spmv_vv4", 0), ("spmv_vv3", 0), ("spmv_vv3", 1), ("spmv_ss", 0), ("spmv_sv", 0), ("spmv_sv", 1), ("spmv_vs", 0), ("spmv_vs", 1), ("multi_dot", 8),
                ("multi_dot", 64), ("cgs_update_scale", 8), ("cgs_update_scale", 64), ("blas1", 0), ("scale_val", 0), ("depart", 0)]
        for name, k in plan:
            ms, by = be.op_bench(name, k=k, reps=a.reps)
            print(json.dumps(dict(kernel=name, k=k, ms=ms, MB=by / 1e6, GBps=by / 1e6 / ms, frac=by / 1e6 / ms / pk)), flush=True)
    else:
        ls_type, RI, GM, CG = P.LS_SETTINGS[a.ls]
        be.timer_start()
        X, info = be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case["incL"], case["res"], fetch=False)
        ms = be.timer_stop()
        print(json.dumps(dict(solve_ms=ms, info=info, launches=be.launch_count())))
    be.close()


if __name__ == "__main__":
    main()
