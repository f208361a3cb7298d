#!/usr/bin/env python
"""One Newton-iteration hot path of the bench workload (P10, LS NS) between cudaProfilerStart / Stop, for the ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
        python tools/step_launches.py
    python tools/step_launches.py --summarise gpurun_out/launches.csv > profiles/<name>_summary.json

ncu's per-launch times are cold-cache and serialised: the kernels' SHARES of the step are what is compared with the live CUDA-event
shares of the bench line."""
import argparse
import csv
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(dims):
    from svfsiplus_b200 import backend as B
    from svfsiplus_b200 import partition as PT
    from svfsiplus_b200 import problem as P
    part, be = PT.setup_distributed_case(dims, 0, 1, 0)
    tDof = part["Ag"].shape[1]
    props = B.fluid_props(tDof=tDof, **part["props"])
    ls_type, RI, GM, CG = P.LS_SETTINGS["NS"]

    def step():
        be.zero(4)
        be.assemble_fluid(props)
        return be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, part["incL"], part["res"], fetch=False)[1]

    be.state_set(tDof, part["Ag"], part["Yg"], part["Bf"])
    step()
    rt = ctypes.CDLL("libcudart.so")
    l0 = be.launch_count()
    rt.cudaProfilerStart()
    info = step()
    rt.cudaProfilerStop()
    print(json.dumps({"launches": be.launch_count() - l0, "itr": [info["RI"]["itr"], info["GM"]["itr"], info["CG"]["itr"]]}))
    be.close()


def summarise(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("svb200::", "")
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)      # -> ms
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = {"launches": sum(a[0] for a in agg.values()), "sum_ms": tot,
           "kernels": {k: {"launches": a[0], "ms": round(a[1], 3), "share": round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[96, 96, 181])
    ap.add_argument("--summarise")
    a = ap.parse_args()
    summarise(a.summarise) if a.summarise else run(tuple(a.dims))
