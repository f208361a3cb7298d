"""Worker of tests/test_multigpu.py, launched with torchrun (one process per GPU, NCCL).

Every rank cuts the same global pipe case into slabs, sets its backend up exactly like a multi-rank
svMultiPhysics run would (local CSR, FSILS node reordering and overlap lists, NCCL communicator),
runs assembly + commu(R) + solve, and rank 0 compares the gathered solution with the compiled
reference's single-rank solution (the standard the reference applies to its own 1/3/4-rank runs,
tests/conftest.py) and, when available, with the reference run on threads-as-ranks on the SAME partition.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from svfsiplus_b200 import backend as B  # noqa: E402
from svfsiplus_b200 import partition as PT  # noqa: E402
from svfsiplus_b200 import problem as P  # noqa: E402


def main():
    dims = tuple(int(v) for v in sys.argv[1:4])
    ls_name = sys.argv[4]
    partition = sys.argv[5] if len(sys.argv) > 5 else "slab"      # slab | rcb (irregular: more neighbours per rank, unbalanced halos)
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()

    case = P.pipe_case(*dims)
    if partition == "rcb":
        parts = PT.split_case(case, world, PT.element_partition_rcb(case["mesh"], world))
    elif partition == "metis":
        parts = PT.split_case(case, world, PT.element_partition_metis(case["mesh"], world))
    else:
        parts = PT.split_case(case, world)
    shared = PT.face_shared_flags(parts)
    part = parts[rank]
    part["face_shared"] = shared
    layout = PT.lhs_layout(rank, [p["gNodes"] for p in parts], part["gnNo"])
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t = torch.from_numpy(B.unique_id().copy()).cuda()
    dist.broadcast(t, src=0)
    be = PT.setup_rank_backend(part, layout, local, t.cpu().numpy())

    transport = be.comm_transport()
    be.state_set(part["Ag"].shape[1], part["Ag"], part["Yg"], part["Bf"])
    be.zero(4)
    be.assemble_fluid(B.fluid_props(tDof=part["Ag"].shape[1], **part["props"]))
    Rloc, Vloc = be.get_R(), be.get_Val()              # partial (un-summed) like com_mod.R before commu
    be.commu_R()
    Rsum = be.get_R()
    ls_type, RI, GM, CG = P.LS_SETTINGS[ls_name]
    X, info = be.solve(ls_type, B.PREC_FSILS, RI, GM, CG, part["incL"], part["res"])

    # gather to rank 0
    objs = [None] * world
    dist.all_gather_object(objs, dict(g=part["gNodes"], X=X, R=Rsum, Rloc=Rloc, Vloc=Vloc, info=info, transport=transport,
                                      nreq=len(layout["reqs"])))
    ok = True
    report = {}
    if rank == 0:
        nNo = case["mesh"].nNo
        Xg = np.zeros((nNo, 4)); Rg = np.zeros((nNo, 4))
        for o in objs:
            Xg[o["g"]] = o["X"]; Rg[o["g"]] = o["R"]
        report["itr"] = [o["info"]["RI"]["itr"] for o in objs]
        report["gm_itr"] = [o["info"]["GM"]["itr"] for o in objs]
        report["cg_itr"] = [o["info"]["CG"]["itr"] for o in objs]
        report["transport"] = [o["transport"] for o in objs]
        report["neighbours"] = [o["nreq"] for o in objs]
        report["X_max"] = float(np.abs(Xg).max())
        report["suc"] = [o["info"]["RI"]["suc"] for o in objs]
        # overlap nodes hold identical values on both owners
        for o in objs:
            report.setdefault("overlap_X", []).append(float(np.abs(Xg[o["g"]] - o["X"]).max()))
        # the same global system on ONE GPU (rank 0's device): partition-independence of the product itself
        be1 = P.setup_backend(case, device=local)
        X1, info1 = P.newton_linear_step(be1, case, ls=ls_name)
        be1.close()
        report["X_vs_1gpu"] = float(np.linalg.norm(Xg - X1) / np.linalg.norm(X1))
        report["itr_1gpu"] = [int(info1["RI"]["itr"]), int(info1["GM"]["itr"]), int(info1["CG"]["itr"])]
        from oracle import ref, refcase
        report["oracle"] = bool(ref.available())
        if ref.available():
            Rr, Vr, Xr, oref = refcase.reference_step(case, ls_name)
            report["R_vs_1rank"] = float(np.abs(Rg - Rr).max() / np.abs(Rr).max())
            report["X_vs_1rank"] = float(np.linalg.norm(Xg - Xr) / np.linalg.norm(Xr))
            report["itr_1rank"] = int(oref["itr"])
            # the reference on the same partition (threads as ranks): local systems from OUR assembly
            rr = ref.RefRanks([dict(gnNo=p["gnNo"], gNodes=p["gNodes"], rowPtr=p["rowPtr"], colPtr=p["colPtr"],
                                    faces=[dict(nodes=f["nodes"], dof=f["dof"], bGrp=f["bGrp"], val=f["val"]) for f in p["faces"]])
                               for p in parts])
            Rs = rr.commuv(4, [o["Rloc"] for o in objs])
            report["commu_R"] = max(float(np.abs(Rs[i] - objs[i]["R"]).max() / np.abs(Rr).max()) for i in range(world))
            Xs, _, outs = rr.solve(4, refcase._ls_vector(P.LS_SETTINGS[ls_name]), 701, Rs, [o["Vloc"] for o in objs],
                                   case["incL"], case["res"])
            Xm = np.zeros((nNo, 4))
            for i, p in enumerate(parts):
                Xm[p["gNodes"]] = Xs[i]
            report["X_vs_Nrank_ref"] = float(np.linalg.norm(Xg - Xm) / np.linalg.norm(Xm))
            report["itr_Nrank_ref"] = int(outs[0]["itr"])
            rr.close()
        print("MULTIGPU_REPORT " + json.dumps(report), flush=True)
    be.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
