#!/usr/bin/env python
"""bench.py — Newton-iteration throughput of the nonlinear-step hot path (assembly + FSILS-equivalent
solve) on the synthetic 10M-tet pipe (P10), one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU code (oracle/_ref)

A "step" is one Newton iteration's hot path: ls_alloc (zero R/Val) + construct_fluid over the whole
mesh + fsils_solve with the <LS> block of tests/cases/fluid/pipe_RCR_3d/solver.xml.
  value : Newton iterations / s, inputs (Ag, Yg, Bf) resident in HBM when the timed region starts
  e2e   : the same metric through the C ABI with HOST buffers: every step copies Ag/Yg/Bf host->device
          from pinned memory and the solution device->host inside the timed region
Timing: CUDA events on the library's launch stream (b200_timer), barrier + device synchronize on both
sides, max over ranks.  The matrix (3.2 GB at P10) is far larger than L2 (126 MB), so no L2 flush is
needed between iterations (stated in config.l2).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NS Newton-iters/s, 10M-tet pipe (assembly+GMRES)"
UNIT = "Newton-iters/s"
P10 = (96, 96, 181)


# dram__bytes_read.sum + dram__bytes_write.sum per launch at P10 on one GPU, from the `ncu --set full` capture
# profiles/r01_tour_c_ncu_raw.csv (same kernels, same data set-up as the bench)
NCU_TRAFFIC_P10 = {"spmv_vv4": 3.5201e9 + 0.0319e9, "spmv_vv3": 2.0842e9 + 0.0425e9, "spmv_sv": 0.7326e9 + 0.0516e9,
                   "spmv_vs": 1.0476e9 + 0.0084e9, "multi_dot": None, "cgs_update_scale": None}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows = []
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _ref_ranks(args, dims):
    """Ranks (= host threads) of the reference arm: its parallelism is MPI ranks; no MPI runtime exists on the box,
    so the ranks are threads of the in-process MPI stand-in (oracle/mpi_stub).  At least two hex layers per rank."""
    if args.ref_ranks > 0:
        return args.ref_ranks
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    return max(1, min(ncpu, 16, dims[2] // 2))


def _reference_sample(args, ls_name):
    from oracle import ref, refcase
    from svfsiplus_b200 import problem as P
    dims = tuple(args.ref_dims)
    case = P.pipe_case(*dims)
    ntet = case["mesh"].nEl
    scale = ntet / float(6 * P10[0] * P10[1] * P10[2])
    nr = _ref_ranks(args, dims)
    r = refcase.reference_step_ranks(case, ls_name, nr)
    return case, dims, ntet, scale, nr, r


def _sample_text(dims, ntet, scale, nr, r, ls_name):
    return (f"pipe {dims[0]}x{dims[1]}x{dims[2]} = {ntet} tets ({100*scale:.2f}% of P10) on {nr} rank(s) = host threads of the "
            f"in-process MPI stand-in: construct_fluid ({r['asm_s']:.2f} s, slowest rank) + commu + fsils_solve {ls_name} "
            f"({r['solve_s']:.2f} s, itr {r['itr']}/{r['GM_itr']}/{r['CG_itr']}); iters/s scaled by the tet ratio to the 10M-tet unit "
            f"(optimistic for the CPU: Krylov counts grow with refinement)")


def run_reference(args):
    """The reference's own CPU implementation (compiled from its sources, oracle/_ref) on a bounded sample of the
    workload, on as many host threads as it can use (one rank per thread, see _ref_ranks)."""
    rank, world, _ = _dist()
    if rank != 0:
        return
    from oracle import ref
    times = []
    r = None
    for i in range(args.warmup + args.steps):
        case, dims, ntet, scale, nr, r = _reference_sample(args, args.ls)
        if i >= args.warmup:
            times.append(r["wall_s"])
    ms = 1e3 * float(np.mean(times))
    value = (1.0 / (ms * 1e-3)) * scale
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"P10 pipe 96x96x181 (10,008,576 TET4), NS VMS, LS {args.ls}; reference timed on a bounded sample",
                   "sample_dims": list(dims), "ls": args.ls, "krylov_itr": r["itr"], "gm_itr": r["GM_itr"], "cg_itr": r["CG_itr"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nr, "kind": "reference" if ref.available() else "port",
                         "sample": _sample_text(dims, ntet, scale, nr, r, args.ls)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline_leg(args, ls_name):
    """Bounded CPU sample for the `cpu_baseline` object of the GPU arm (rank 0, N=1 only)."""
    try:
        from oracle import ref
        if not ref.available():
            return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref not present"}
        case, dims, ntet, scale, nr, r = _reference_sample(args, ls_name)
        return {"value": (1.0 / r["wall_s"]) * scale, "unit": UNIT, "cores": nr, "kind": "reference",
                "sample": _sample_text(dims, ntet, scale, nr, r, ls_name),
                "assembly_us_per_tet_per_rank": 1e6 * r["asm_s"] * nr / ntet}
    except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU line
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e}"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from svfsiplus_b200 import backend as B
    from svfsiplus_b200 import problem as P

    rank, world, local = _dist()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    if world > 1:
        from svfsiplus_b200 import partition as PT
        dims = PT.weak_dims(tuple(args.dims), world)      # default --dims = P10 -> configs[2] at N = 8
        case, be = PT.setup_distributed_case(dims, rank, world, local, dist)
    else:
        dims = tuple(args.dims)
        from svfsiplus_b200 import backend as B
        be = B.Backend(local)
        case = P.pipe_case(*dims, pattern=lambda n, ien: be.pattern(n, [ien]))       # lhsa on the device (b200_pattern_*)
        be = P.setup_backend(case, device=local, be=be)
    nNo_local = be.nNo
    tDof = case["Ag"].shape[1]
    ntet_total = 6 * dims[0] * dims[1] * dims[2]

    # pinned host buffers for the end-to-end leg
    pin = {k: torch.from_numpy(np.ascontiguousarray(case[k])).pin_memory() for k in ("Ag", "Yg", "Bf")}
    out_pin = torch.empty((nNo_local, 4), dtype=torch.float64).pin_memory()
    ls_type, RI, GM, CG = P.LS_SETTINGS[args.ls]
    props = B.fluid_props(tDof=tDof, **case["props"])

    def step_resident():
        be.zero(4)
        be.assemble_fluid(props)
        if world > 1:
            be.commu_R()
        _, info = be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case["incL"], case["res"], fetch=False)
        return info

    def step_e2e():
        be.state_set(tDof, pin["Ag"].data_ptr(), pin["Yg"].data_ptr(), pin["Bf"].data_ptr())
        be.zero(4)
        be.assemble_fluid(props)
        if world > 1:
            be.commu_R()
        _, info = be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case["incL"], case["res"], out=out_pin.data_ptr(), fetch=True)
        return info

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    be.state_set(tDof, pin["Ag"].data_ptr(), pin["Yg"].data_ptr(), pin["Bf"].data_ptr())
    info = None
    for _ in range(args.warmup):
        info = step_resident()

    # profiled pass (untimed, after the warm-up): a CUDA-event pair around every kernel class on the
    # launch stream gives the shares of the step and names the dominant kernel class
    be.profile(True)
    barrier()
    be.timer_start()
    for _ in range(args.prof_steps):
        step_resident()
    prof_ms = be.timer_stop()
    barrier()
    prof = be.profile_read()
    be.profile(False)
    dom = max(prof, key=lambda k: prof[k]["ms"])

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timed region: only the dominant class records events (~1.5 k pairs per step instead of ~60 k, which
    # would cost ~4 % of the step), so roofline.achieved is measured live inside the timed region
    be.profile(2 + B.KERNEL_CLASSES.index(dom))
    l0 = be.launch_count()
    barrier()
    be.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        info = step_resident()
    dev_ms = be.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = be.launch_count() - l0
    dom_live = be.profile_read()[dom]
    be.profile(False)
    dev_ms = maxreduce(dev_ms)
    ms_per_step = dev_ms / args.steps

    # end-to-end leg (same K)
    step_e2e()
    barrier()
    be.timer_start()
    for _ in range(args.steps):
        step_e2e()
    e2e_ms = be.timer_stop()
    barrier()
    e2e_ms = maxreduce(e2e_ms) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # stand-alone block SpMV (the north-star kernel): CUDA events over 50 back-to-back launches.  Collective
    # for N > 1 (assembly is followed by commu(R), every product by its overlap add): ALL ranks run it.
    P.assemble(be, case, upload=False)
    spmv_ms, spmv_bytes = be.op_bench("spmv_vv4", reps=50)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # dominant kernel class of the timed region -> roofline
    peak, peak_src = _peaks()
    tot_k = sum(v["ms"] for v in prof.values())
    d = dom_live
    achieved = (d["bytes"] / 1e9) / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0
    shares = {k: round(v["ms"] / prof_ms, 4) for k, v in prof.items() if v["ms"] > 0}
    per_class = {k: {"ms_per_launch": v["ms"] / max(v["launches"], 1), "launches_per_step": v["launches"] / args.prof_steps,
                     "GBps": (v["bytes"] / 1e9) / (v["ms"] * 1e-3) if v["ms"] > 0 else None}
                 for k, v in prof.items() if v["launches"] > 0}
    nnz, nNo = be.nnz, be.nNo
    spmv_gbs = spmv_bytes / 1e9 / (spmv_ms * 1e-3)

    scale = ntet_total / float(6 * P10[0] * P10[1] * P10[2])
    value = (1e3 / ms_per_step) * (scale if world > 1 else 1.0)
    e2e_value = (1e3 / e2e_ms) * (scale if world > 1 else 1.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pipe {dims[0]}x{dims[1]}x{dims[2]} = {ntet_total} TET4, NS VMS P1-P1, one Newton iteration "
                               f"(ls_alloc + construct_fluid + fsils_solve LS {args.ls}, pipe_RCR_3d solver.xml parameters)",
                   "ls": args.ls, "nNo": int(nNo), "nnz_blocks": int(nnz), "parallelism": f"dd{world}",
                   "l2": "inputs larger than L2 (Val 3.2 GB vs 126 MB L2): no flush between iterations",
                   "krylov_itr": info["RI"]["itr"], "gm_itr": info["GM"]["itr"], "cg_itr": info["CG"]["itr"],
                   "suc": info["RI"]["suc"], "wall_ms_per_step": wall_ms / args.steps,
                   "value_note": "N>1: Newton-iters/s x (total tets / 10,008,576), i.e. 10M-tet-equivalent iterations/s"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int((2 * tDof + 3) * nNo_local * 8),
                "d2h_bytes_per_step": int(4 * nNo_local * 8), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (NCU_TRAFFIC_P10.get(dom) if (world == 1 and tuple(dims) == P10) else None),
                     "traffic_source": "ncu --set full, profiles/r01_tour_c_ncu_raw.csv (per launch)", "peak_source": peak_src,
                     "share_of_step": d["ms"] / dev_ms, "launches": d["launches"],
                     "bytes_per_launch": d["bytes"] / max(d["launches"], 1)},
        "kernel_shares": shares, "kernels": per_class, "kernel_time_frac_of_step": tot_k / prof_ms, "profiled_ms_per_step": prof_ms / args.prof_steps,
        "spmv": {"kernel": "k_spmv_vv4", "ms": spmv_ms, "bytes": spmv_bytes, "GBps": spmv_gbs, "frac_of_peak": spmv_gbs / peak,
                 "frac_of_8TBps": spmv_gbs / 8000.0},
        "clocks": clocks,
    }
    # Refinement-independent work rate.  Krylov iteration counts grow with refinement (the reference algorithm has no multigrid),
    # so the weak-scaled `value` at N > 1 contains that growth as well as the communication cost; tets x inner Krylov iterations
    # per second separates the two (same definition at every N; the driver's efficiency is computed from `value`, not from this).
    try:
        inner = int(info["GM"]["itr"]) + int(info["CG"]["itr"])
        if inner <= 0:
            inner = int(info["RI"]["itr"])
        line["krylov_work_rate"] = {"value": ntet_total * inner / (ms_per_step * 1e-3), "unit": "tet x inner Krylov iterations / s",
                                    "inner_iterations_per_step": inner, "per_gpu": ntet_total * inner / (ms_per_step * 1e-3) / world}
    except Exception:                      # never let a reporting extra cost the bench line
        pass
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(args, args.ls)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # NCCL prints a "NCCL version ..." banner to STDOUT at NCCL_DEBUG=VERSION/WARN; stdout must carry one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = "NONE"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ls", default="NS", choices=["NS", "GMRES", "BICGS"])
    ap.add_argument("--dims", type=int, nargs=3, default=list(P10), help="pipe hex counts nx ny nz (default P10)")
    ap.add_argument("--ref-dims", type=int, nargs=3, default=[24, 24, 48], help="bounded CPU sample of the workload")
    ap.add_argument("--ref-ranks", type=int, default=0, help="ranks (threads) of the reference arm; 0 = min(host cores, 16, layers/2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prof-steps", type=int, default=1, help="extra steps run with per-kernel CUDA events (shares, roofline)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
