#!/usr/bin/env python
"""bench.py — Newton-iteration throughput of the nonlinear-step hot path (assembly + FSILS-equivalent
solve) on the synthetic 10M-tet pipe (P10), one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU code (oracle/_ref)
    python bench.py --workload struct_block|ustruct_block|fsi_pipe      # BASELINE.json configs[3] / configs[4]

A "step" is one Newton iteration's hot path: ls_alloc (zero R/Val) + construct_fluid over the whole
mesh + fsils_solve with the <LS> block of tests/cases/fluid/pipe_RCR_3d/solver.xml.
  value : Newton iterations / s of the 10M-tet pipe on N GPUs (N > 1: the same mesh split over the ranks - BASELINE's metric reads
          "10M-tet pipe @1-8 B200", scaling "strong"), inputs (Ag, Yg, Bf) resident in HBM when the timed region starts
  e2e   : the same metric through the C ABI with HOST buffers: every step copies Ag/Yg/Bf host->device
          from pinned memory and the solution device->host inside the timed region
  weak  : (N > 1) the pipe refined to N x the tets (configs[2]; 80 M at N = 8), 10M-tet-equivalent iterations/s
  fixed_work, same_config, golden_p10: see DESIGN.md section 6
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
# LS NS with fixed iteration counts (see `fixed_work` in run_gpu); 796 = LS_NS (the reference's code, liner_solver/fils_struct.hpp)
FIXED_WORK_LS = (796, (0.0, 0.0, 2, 250), (0.0, 0.0, 1, 50), (0.0, 0.0, 200, 0))


# dram__bytes_read.sum + dram__bytes_write.sum per launch at P10 on one GPU, from `ncu --set full` captures of the same kernels on the
# same data set-up as the bench: profiles/r01_tour_c_ncu_raw.csv (round 1) and profiles/r02_vv3_variants_ncu_raw.csv (the column-owner
# dof-3 kernel that is the default since round 2).  The Arnoldi kernels' traffic depends on the basis depth of the launch, so there is
# no single per-launch figure for them (null; at a depth of 64 ncu reads 2.68 GB against 2.71 GB algorithmic).
NCU_TRAFFIC_P10 = {"spmv_vv4": 3.5201e9 + 0.0319e9, "spmv_vv3": 2.0080e9 + 0.0430e9, "spmv_sv": 0.7326e9 + 0.0516e9,
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


def workload_dims(args, world):
    """The pipe the line is quoted on: the 10M-tet pipe P10 (BASELINE.json: "10M-tet pipe (assembly+GMRES) @1-8 B200") whatever N
    is - N GPUs split the SAME mesh (strong scaling).  The refined pipes of configs[2] (N x the tets on N GPUs) are timed in the
    same run and reported in the `weak` object."""
    return tuple(args.dims)


def make_config(args, world):
    """`config` is the SAME object in both arms (the reference arm times a bounded sample of this workload and says so in
    cpu_baseline.sample; run-dependent numbers such as Krylov counts live under `run`, not here)."""
    dims = workload_dims(args, world)
    ntet = 6 * dims[0] * dims[1] * dims[2]
    return {"workload": f"pipe {dims[0]}x{dims[1]}x{dims[2]} = {ntet} TET4, NS VMS P1-P1, one Newton iteration "
                        f"(ls_alloc + construct_fluid + fsils_solve LS {args.ls}, pipe_RCR_3d solver.xml parameters)",
            "ls": args.ls, "parallelism": f"dd{world}",
            "l2": "inputs larger than L2 (Val 3.2 GB vs 126 MB L2): no flush between iterations",
            "value_note": "Newton iterations per second of THIS mesh on N GPUs (strong scaling); the `weak` object holds the pipe "
                          "refined to N x the tets, in 10M-tet-equivalent iterations/s"}


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


_REF_CASES = {}


def _reference_sample(args, ls_name):
    """One Newton-iteration hot path of the compiled reference on the bounded sample --ref-dims (the case is built once)."""
    from oracle import refcase
    from svfsiplus_b200 import problem as P
    dims = tuple(args.ref_dims)
    if dims not in _REF_CASES:
        _REF_CASES[dims] = P.pipe_case(*dims)
    case = _REF_CASES[dims]
    ntet = case["mesh"].nEl
    scale = ntet / float(6 * P10[0] * P10[1] * P10[2])
    nr = _ref_ranks(args, dims)
    r = refcase.reference_step_ranks(case, ls_name, nr)
    return case, dims, ntet, scale, nr, r


def _ref_build():
    return "-O3 build (oracle/_ref/o3)" if "o3" in os.environ.get("SVREF_LIB", "") else "-O2 build"


def _sample_text(dims, ntet, scale, nr, r, ls_name):
    return (f"compiled reference, {_ref_build()}; pipe {dims[0]}x{dims[1]}x{dims[2]} = {ntet} tets ({100*scale:.2f}% of P10) on {nr} rank(s) = host threads of the "
            f"in-process MPI stand-in: construct_fluid ({r['asm_s']:.2f} s, slowest rank) + commu + fsils_solve {ls_name} "
            f"({r['solve_s']:.2f} s, itr {r['itr']}/{r['GM_itr']}/{r['CG_itr']}); iters/s scaled by the tet ratio to the 10M-tet unit "
            f"(EXTRAPOLATED, optimistic for the CPU: Krylov counts grow with refinement; the measured same-size ratio is the GPU "
            f"line's `same_config` object)")


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
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": make_config(args, max(1, args.gpus)),
        "run": {"sample_dims": list(dims), "sample_tets": int(ntet), "extrapolated": True, "same_config": False,
                "sample_iters_per_s": 1.0 / (ms * 1e-3),
                "krylov_itr": r["itr"], "gm_itr": r["GM_itr"], "cg_itr": r["CG_itr"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nr, "kind": "reference" if ref.available() else "port",
                         "sample": _sample_text(dims, ntet, scale, nr, r, args.ls), "extrapolated": True},
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
                "sample": _sample_text(dims, ntet, scale, nr, r, ls_name), "extrapolated": True,
                "assembly_us_per_tet_per_rank": 1e6 * r["asm_s"] * nr / ntet,
                "_sample": {"dims": list(dims), "tets": int(ntet), "iters_per_s": 1.0 / r["wall_s"], "wall_s": r["wall_s"],
                            "krylov_itr": int(r["itr"]), "gm_itr": int(r["GM_itr"]), "cg_itr": int(r["CG_itr"])}}
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

    from svfsiplus_b200 import partition as PT
    dims = workload_dims(args, world)                     # default --dims = P10 -> configs[2] at N = 8

    def build(dims_):
        # the pipe is composed of 8 generation blocks whatever N is (partition.local_slab_case): 1, 2, 4 and 8 ranks hold the SAME
        # global mesh and state, so the strong-scaling object and the one-GPU line solve identical systems
        return PT.setup_distributed_case(dims_, rank, world, local, dist)

    case, be = build(dims)
    nNo_local = be.nNo
    tDof = case["Ag"].shape[1]
    ntet_total = 6 * dims[0] * dims[1] * dims[2]
    transport = be.comm_transport()

    # pinned host buffers for the end-to-end leg
    def pinned(c, b):
        pin_ = {k: torch.from_numpy(np.ascontiguousarray(c[k])).pin_memory() for k in ("Ag", "Yg", "Bf")}
        return pin_, torch.empty((b.nNo, 4), dtype=torch.float64).pin_memory()

    pin, out_pin = pinned(case, be)
    LS = P.LS_SETTINGS[args.ls]
    props = B.fluid_props(tDof=tDof, **case["props"])

    def make_steps(be_, case_, pin_, out_pin_, ls):
        ls_type, RI, GM, CG = ls
        props_ = B.fluid_props(tDof=case_["Ag"].shape[1], **case_["props"])
        tD = case_["Ag"].shape[1]

        def resident():
            be_.zero(4)
            be_.assemble_fluid(props_)
            if world > 1:
                be_.commu_R()
            _, info_ = be_.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case_["incL"], case_["res"], fetch=False)
            return info_

        def e2e():
            be_.state_set(tD, pin_["Ag"].data_ptr(), pin_["Yg"].data_ptr(), pin_["Bf"].data_ptr())
            be_.zero(4)
            be_.assemble_fluid(props_)
            if world > 1:
                be_.commu_R()
            _, info_ = be_.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case_["incL"], case_["res"], out=out_pin_.data_ptr(), fetch=True)
            return info_
        return resident, e2e

    step_resident, step_e2e = make_steps(be, case, pin, out_pin, LS)

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

    def global_norms(xpin, case_):
        """||X(:, j)|| over the global nodes, every node counted once (an interface plane belongs to the lower rank), and the
        solution at the golden file's probe nodes (zeros where another rank owns the node; summed over ranks)."""
        X = xpin.numpy()
        nx_, ny_, _ = case_["mesh"].shape
        plane_ = (nx_ + 1) * (ny_ + 1)
        own = X.shape[0] - (plane_ if rank < world - 1 else 0)
        v = torch.tensor((X[:own] ** 2).sum(axis=0), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v)
        return [float(t) for t in v.sqrt().cpu()]

    x_norm = global_norms(out_pin, case)

    # stand-alone block SpMV (the north-star kernel): CUDA events over 50 back-to-back launches.  Collective
    # for N > 1 (assembly is followed by commu(R), every product by its overlap add): ALL ranks run it.
    P.assemble(be, case, upload=False)
    spmv_ms, spmv_bytes = be.op_bench("spmv_vv4", reps=50)
    barrier()

    def timed(fn, n):
        barrier()
        be_timer.timer_start()
        inf = None
        for _ in range(n):
            inf = fn()
        ms = be_timer.timer_stop()
        barrier()
        return maxreduce(ms) / n, inf

    def counts(inf):
        return {"krylov_itr": inf["RI"]["itr"], "gm_itr": inf["GM"]["itr"], "cg_itr": inf["CG"]["itr"], "suc": inf["RI"]["suc"]}

    # (1) fixed work: the SAME partitioned mesh, LS NS with every tolerance 0 and fixed iteration limits, so every N does
    # identical Krylov work per node (2 outer iterations x [2 x 50 GMRES + 200 Schur-CG]); the ratio of this number between
    # N and 1 is pure kernel + communication cost, free of the refinement-driven growth of the iteration counts.
    be_timer = be
    fw_steps = max(2, min(args.steps, 5))
    fw_res, _ = make_steps(be, case, pin, out_pin, FIXED_WORK_LS)
    fw_res()
    fw_ms, fw_info = timed(fw_res, fw_steps)
    fw_inner = int(fw_info["GM"]["itr"]) + int(fw_info["CG"]["itr"])
    fixed_work = {"ls": "NS, relTol = absTol = 0, RI mItr 2, GM 1 x 50, CG 200 (fixed iteration counts)", "steps": fw_steps,
                  "ms_per_step": fw_ms, **counts(fw_info), "tets": ntet_total,
                  "work_rate": ntet_total * fw_inner / (fw_ms * 1e-3), "work_rate_per_gpu": ntet_total * fw_inner / (fw_ms * 1e-3) / world,
                  "unit": "tet x inner Krylov iterations / s", "scaling": "strong (the headline mesh split over N ranks); the weak object carries its own fixed-work time"}
    be.close()

    # (2) weak scaling (BASELINE.json configs[2]): the same pipe refined to N x the tets (192 x 192 x 362 = 80 M at N = 8), each rank
    # its z-slab; value in 10M-tet-equivalent iterations/s.  Krylov counts grow with refinement (the reference algorithm has no
    # multilevel preconditioner), which this number contains together with the communication cost; `fixed_work` separates them.
    weak = None
    if world > 1 and not args.no_weak:
        wdims = PT.weak_dims(tuple(args.dims), world)
        case_s, be_s = build(wdims)
        pin_s, out_s = pinned(case_s, be_s)
        be_timer = be_s
        s_res, s_e2e = make_steps(be_s, case_s, pin_s, out_s, LS)
        be_s.state_set(case_s["Ag"].shape[1], pin_s["Ag"].data_ptr(), pin_s["Yg"].data_ptr(), pin_s["Bf"].data_ptr())
        for _ in range(2):
            s_res()
        s_steps = max(2, min(args.steps, 3))
        s_ms, s_info = timed(s_res, s_steps)
        s_e2e()
        s_e2e_ms, _ = timed(s_e2e, s_steps)
        fw_s, _ = make_steps(be_s, case_s, pin_s, out_s, FIXED_WORK_LS)
        fw_s()
        s_fw_ms, s_fw_info = timed(fw_s, s_steps)
        wtets = 6 * wdims[0] * wdims[1] * wdims[2]
        wscale = wtets / float(6 * P10[0] * P10[1] * P10[2])
        w_inner = int(s_fw_info["GM"]["itr"]) + int(s_fw_info["CG"]["itr"])
        weak = {"dims": list(wdims), "tets": wtets, "steps": s_steps, "ms_per_step": s_ms, "value": (1e3 / s_ms) * wscale,
                "e2e_value": (1e3 / s_e2e_ms) * wscale, "unit": UNIT + " x (tets / 10,008,576)", **counts(s_info),
                "fixed_work_ms_per_step": s_fw_ms, "fixed_work_inner": w_inner,
                "fixed_work_rate_per_gpu": wtets * w_inner / (s_fw_ms * 1e-3) / world,
                "krylov_work_rate_per_gpu": wtets * (int(s_info["GM"]["itr"]) + int(s_info["CG"]["itr"])) / (s_ms * 1e-3) / world}
        be_s.close()

    # (3) same-config measurement against the reference (N = 1): the bounded sample the reference arm times, on the GPU
    same_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        sd = tuple(args.ref_dims)
        case_c = P.pipe_case(*sd)
        be_c = P.setup_backend(case_c, device=local)
        pin_c, out_c = pinned(case_c, be_c)
        be_timer = be_c
        c_res, c_e2e = make_steps(be_c, case_c, pin_c, out_c, LS)
        be_c.state_set(case_c["Ag"].shape[1], pin_c["Ag"].data_ptr(), pin_c["Yg"].data_ptr(), pin_c["Bf"].data_ptr())
        for _ in range(3):
            c_res()
        c_ms, c_info = timed(c_res, max(args.steps, 5))
        c_e2e()
        c_e2e_ms, _ = timed(c_e2e, max(args.steps, 5))
        same_gpu = {"dims": list(sd), "tets": int(case_c["mesh"].nEl), "gpu_ms_per_step": c_ms, "gpu_value": 1e3 / c_ms,
                    "gpu_e2e_value": 1e3 / c_e2e_ms, "gpu_counts": counts(c_info)}
        be_c.close()

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

    value = 1e3 / ms_per_step
    e2e_value = 1e3 / e2e_ms
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": make_config(args, world),
        "run": {"nNo": int(nNo), "nnz_blocks": int(nnz), "krylov_itr": info["RI"]["itr"], "gm_itr": info["GM"]["itr"],
                "cg_itr": info["CG"]["itr"], "suc": info["RI"]["suc"], "wall_ms_per_step": wall_ms / args.steps,
                "transport": transport, "X_norm": x_norm},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int((2 * tDof + 3) * nNo_local * 8),
                "d2h_bytes_per_step": int(4 * nNo_local * 8), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (NCU_TRAFFIC_P10.get(dom) if (world == 1 and tuple(dims) == P10) else None),
                     "traffic_source": "ncu --set full, profiles/r01_tour_c_ncu_raw.csv / r02_vv3_variants_ncu_raw.csv (per launch; null for kernels whose traffic depends on the basis depth)", "peak_source": peak_src,
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
    line["fixed_work"] = fixed_work
    if weak is not None:
        line["weak"] = weak
    # benchmark-size parity: counts and solution norms of the compiled reference on the same P10 system (tests/golden/p10_<ls>_counts.json,
    # generated offline by tests/golden/make_golden_p10.py); a plain file read, nothing of oracle/ is executed here
    try:
        gp = os.path.join(ROOT, "tests", "golden", f"p10_{args.ls.lower()}_counts.json")       # NS (headline) or the plain GMRES variant
        if os.path.exists(gp):
            g = json.load(open(gp))

            def chk(cnt, xn):
                return {"reference": {k: g[k] for k in ("itr", "GM_itr", "CG_itr")},
                        "counts_within_1": bool(abs(cnt["krylov_itr"] - g["itr"]) <= 1),
                        "inner_counts_rel": [abs(cnt["gm_itr"] - g["GM_itr"]) / max(g["GM_itr"], 1), abs(cnt["cg_itr"] - g["CG_itr"]) / max(g["CG_itr"], 1)],
                        "X_norm_rel": [abs(a - b) / b for a, b in zip(xn, g["X_norm"])]}
            if list(dims) == g["dims"]:            # every N solves the same global system: same counts, same solution norms
                line["golden_p10"] = chk({"krylov_itr": info["RI"]["itr"], "gm_itr": info["GM"]["itr"], "cg_itr": info["CG"]["itr"]}, x_norm)
    except Exception as e:
        line["golden_p10"] = {"error": str(e)}
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline_leg(args, args.ls)
        smp = cb.pop("_sample", None)
        line["cpu_baseline"] = cb
        if smp and same_gpu:
            # the one ratio in this record that is a MEASUREMENT on identical inputs (same mesh, same <LS> block, both arms
            # converge their own Krylov loops); the P10 ratio the driver computes from `value` is an extrapolation of this sample
            line["same_config"] = {**same_gpu, "ref_value": smp["iters_per_s"], "ref_wall_s": smp["wall_s"], "ref_cores": cb["cores"],
                                   "ref_counts": {"krylov_itr": smp["krylov_itr"], "gm_itr": smp["gm_itr"], "cg_itr": smp["cg_itr"]},
                                   "unit": "Newton-iters/s on this mesh", "ratio": same_gpu["gpu_value"] * smp["wall_s"],
                                   "e2e_ratio": same_gpu["gpu_e2e_value"] * smp["wall_s"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------------
# the other GPU configurations of BASELINE.json: configs[3] (struct / ustruct hyperelastic block compression) and configs[4]
# (FSI pipe: fluid + struct domains in one block system).  One GPU; under torchrun every rank runs an independent replica
# ("replicas only": these lines have no partitioned variant yet) and the value is the sum.
# ---------------------------------------------------------------------------------------------------------------------------
WORKLOADS = {
    # name: (description, default size)
    "struct_block": ("struct equation, neo-Hookean ST91 block compression (tests/cases/struct/block_compression/solver.xml: "
                     "BICG 1e-12 / 600), HEX8 n^3", 160),
    "ustruct_block": ("ustruct equation, P1-P1 VMS neo-Hookean block compression (tests/cases/ustruct/block_compression/P1P1_VMS/"
                      "solver.xml: GMRES), TET4 6 n^3, with ustruct_r", 100),
    "fsi_pipe": ("FSI equation of tests/cases/fsi/pipe_3d scaled up: fluid lumen on the ALE configuration + neo-Hookean wall in one "
                 "dof-4 block system (GMRES 1e-12 / 100 x 50), TET4 pipe", 0),
}


def run_gpu_workload(args):
    import torch
    from svfsiplus_b200 import backend as B
    from svfsiplus_b200 import problem as P

    rank, world, local = _dist()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = B.Backend(local)
    pat = lambda n, ien: be.pattern(n, [ien])                         # lhsa on the device (b200_pattern_*)
    wl = args.workload
    n = args.size if args.size > 0 else WORKLOADS[wl][1]
    t_setup = time.perf_counter()
    if wl == "struct_block":
        case = P.block_case(n, elem="hex", kind="struct", pattern=pat)
        ls_name, dof = "BICGS_STRUCT", 3
        asm = lambda up: P.assemble_solid(be, case, upload=up)
        ref_step = lambda c: __import__("oracle.refcase", fromlist=["x"]).reference_solid_step(c, ls_name)
        small = lambda: P.block_case(args.ref_size or 20, elem="hex", kind="struct")
        size_txt = f"{n}^3 HEX8"
    elif wl == "ustruct_block":
        case = P.ustruct_case(n, elem="tet", pattern=pat)
        ls_name, dof = "GMRES_USTRUCT", 4
        asm = lambda up: P.assemble_ustruct(be, case, upload=up, with_r=True)
        ref_step = lambda c: __import__("oracle.refcase", fromlist=["x"]).reference_ustruct_step(c, ls_name)
        small = lambda: P.ustruct_case(args.ref_size or 14, elem="tet")
        size_txt = f"6 x {n}^3 TET4"
    else:
        dims = tuple(args.dims)
        case = P.fsi_case(*dims, pattern=pat)
        ls_name, dof = "GMRES_FSI", 4
        asm = lambda up: P.assemble_fsi(be, case, upload=up)
        ref_step = lambda c: __import__("oracle.refcase", fromlist=["x"]).reference_fsi_step(c, ls_name)
        small = lambda: P.fsi_case(12, 12, 24)
        size_txt = f"pipe {dims[0]}x{dims[1]}x{dims[2]} TET4"
    be = P.setup_backend(case, device=local, be=be)
    nEl, nNo = int(case["mesh"].nEl), int(be.nNo)
    t_setup = time.perf_counter() - t_setup
    ls_type, RI, GM, CG = P.LS_SETTINGS[ls_name]
    # pinned host copies of the per-step inputs for the end-to-end leg (the case dict then points at pinned memory)
    h2d = 0
    for k in ("Ag", "Yg", "Dg", "Bf", "Ad"):
        if k in case and case[k] is not None:
            t = torch.from_numpy(np.ascontiguousarray(case[k])).pin_memory()
            case[k] = t.numpy()
            case["_pin_" + k] = t
            h2d += t.numel() * 8
    out_pin = torch.empty((nNo, dof), dtype=torch.float64).pin_memory()

    def step(upload):
        asm(upload)
        _, inf = be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, case["incL"], case["res"], out=out_pin.data_ptr() if upload else None,
                          fetch=upload)
        return inf

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        barrier()
        be.timer_start()
        inf = None
        for _ in range(nsteps):
            inf = fn()
        ms = be.timer_stop()
        barrier()
        return ms / nsteps, inf

    step(True)
    for _ in range(max(args.warmup - 1, 2)):
        step(False)
    be.profile(True)
    prof_ms, _ = timed(lambda: step(False), 1)
    prof = be.profile_read()
    be.profile(False)
    dom = max(prof, key=lambda k: prof[k]["ms"])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    be.profile(2 + B.KERNEL_CLASSES.index(dom))
    l0 = be.launch_count()
    ms_per_step, info = timed(lambda: step(False), args.steps)
    launches = be.launch_count() - l0
    dom_live = be.profile_read()[dom]
    be.profile(False)
    e2e_ms, _ = timed(lambda: step(True), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # assembly alone (whole-mesh element kernel + ordered sums), resident inputs
    asm_ms, _ = timed(lambda: asm(False), max(3, args.steps))
    if world > 1:
        t = torch.tensor([ms_per_step, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    achieved = (dom_live["bytes"] / 1e9) / (dom_live["ms"] * 1e-3) if dom_live["ms"] > 0 else 0.0
    line = {
        "metric": f"Newton-iters/s, {wl}", "value": world * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{wl}: {WORKLOADS[wl][0]}; {size_txt} = {nEl} elements, one Newton iteration (ls_alloc + whole-mesh "
                               f"assembly + fsils_solve LS {ls_name})", "ls": ls_name,
                   "parallelism": "replicas only" if world > 1 else "dd1",
                   "l2": "inputs larger than L2 (Val >> 126 MB): no flush between iterations"},
        "run": {"nEl": nEl, "nNo": nNo, "nnz_blocks": int(be.nnz), "dof": dof, "krylov_itr": info["RI"]["itr"], "suc": info["RI"]["suc"],
                "iNorm": info["RI"]["iNorm"], "fNorm": info["RI"]["fNorm"], "setup_s": t_setup},
        "e2e": {"value": world * 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(nNo * dof * 8),
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "share_of_step": dom_live["ms"] / (ms_per_step * args.steps),
                     "launches": dom_live["launches"], "bytes_per_launch": dom_live["bytes"] / max(dom_live["launches"], 1)},
        "kernel_shares": {k: round(v["ms"] / prof_ms, 4) for k, v in prof.items() if v["ms"] > 0},
        "kernels": {k: {"ms_per_launch": v["ms"] / max(v["launches"], 1), "launches_per_step": v["launches"],
                        "GBps": (v["bytes"] / 1e9) / (v["ms"] * 1e-3) if v["ms"] > 0 else None} for k, v in prof.items() if v["launches"] > 0},
        "assembly": {"ms": asm_ms, "ns_per_element": 1e6 * asm_ms / nEl, "share_of_step": asm_ms / ms_per_step},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        try:
            from oracle import ref
            if ref.available():
                c = small()
                t0 = time.perf_counter()
                out = ref_step(c)
                w = time.perf_counter() - t0
                sc = c["mesh"].nEl / float(nEl)
                line["cpu_baseline"] = {"value": sc / w, "unit": UNIT, "cores": 1, "kind": "reference", "extrapolated": True,
                                        "sample": f"{c['name']}: {c['mesh'].nEl} elements ({100 * sc:.3f}% of the workload), the compiled "
                                                  f"reference's construct_* + fsils_solve {ls_name} on one core, {w:.2f} s, itr "
                                                  f"{int(out[-1]['itr'])}; iters/s scaled by the element ratio (optimistic for the CPU)"}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # NCCL prints its banner / log to STDOUT; stdout must carry one JSON line, so the log is routed to stderr (nothing is
    # suppressed: NCCL_DEBUG keeps whatever level the caller asked for)
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    # timing legs of the reference use its -O3 build when it exists (oracle/Makefile target o3; the parity tests keep the default
    # -O2 library) - the faster of the two, i.e. the conservative denominator
    o3 = os.path.join(ROOT, "oracle", "_ref", "o3", "libsvref.so")
    if os.path.exists(o3) and not os.environ.get("SVREF_LIB"):
        os.environ["SVREF_LIB"] = o3
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ls", default="NS", choices=["NS", "GMRES", "BICGS"])
    ap.add_argument("--dims", type=int, nargs=3, default=list(P10), help="pipe hex counts nx ny nz (default P10)")
    ap.add_argument("--ref-dims", type=int, nargs=3, default=[32, 32, 64], help="bounded CPU sample of the workload")
    ap.add_argument("--ref-ranks", type=int, default=0, help="ranks (threads) of the reference arm; 0 = min(host cores, 16, layers/2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the refined (weak-scaling) pipe of configs[2]")
    ap.add_argument("--workload", default="ns_pipe", choices=["ns_pipe"] + sorted(WORKLOADS),
                    help="ns_pipe = the headline metric (configs[1]/[2]); the others are BASELINE.json configs[3] and configs[4]")
    ap.add_argument("--size", type=int, default=0, help="elements per edge of the block workloads (default: the workload's own)")
    ap.add_argument("--ref-size", type=int, default=0, help="elements per edge of the CPU sample of the block workloads")
    ap.add_argument("--prof-steps", type=int, default=1, help="extra steps run with per-kernel CUDA events (shares, roofline)")
    args = ap.parse_args()
    if args.workload != "ns_pipe":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm is implemented for the headline workload ns_pipe only; "
                              "the other workloads carry their CPU sample in cpu_baseline"}))
            return
        if args.workload == "fsi_pipe" and args.dims == list(P10):
            args.dims = [64, 64, 128]
        run_gpu_workload(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
