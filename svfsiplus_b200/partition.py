"""Domain decomposition for the multi-GPU path: one rank (process, GPU) per partition.

The reference partitions ELEMENTS (ParMETIS on the dual graph, Code/Source/solver/distribute.cpp:1683),
gives every rank a local copy of each node its elements touch, builds the local CSR from the local IEN
(lhsa) and lets fsils_lhs_create (Code/Source/liner_solver/lhs.cpp:57-376) reorder the local nodes —
nodes shared with lower ranks first, interior nodes, nodes shared with higher ranks last — and derive
the pairwise overlap lists lhs.cS[].  Everything here follows that scheme; only the partitioner is
different: a pipe is cut into z-slabs of whole hex layers (contiguous element ranges of the generator),
which is what a graph partitioner returns for this geometry up to rounding.  ParMETIS output depends
on the rank count anyway and only changes summation order.

`lhs_layout` (native, csrc/lhs_layout.hpp behind the C ABI) and `lhs_layout_numpy` (an independent numpy restatement)
reproduce fsils_lhs_create's ordering; both are pinned integer for integer against the compiled reference run on
threads-as-ranks (tests/test_partition.py).
"""
from __future__ import annotations

import numpy as np

from . import mesh as M


# ------------------------------------------------------------------------------------------------
# element partition
# ------------------------------------------------------------------------------------------------
def slab_ranges(nz: int, nparts: int):
    """Hex layers [k0, k1) per part, as even as possible (first parts take the remainder)."""
    base, rem = divmod(nz, nparts)
    out, k = [], 0
    for p in range(nparts):
        n = base + (1 if p < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def element_partition_rcb(mesh: M.Mesh, nparts: int) -> np.ndarray:
    """part[e] for ANY mesh: recursive coordinate bisection of the element centroids (native, b200_partition_rcb) - the
    stand-in for the reference's ParMETIS call (distribute.cpp:1683) where the mesh is not generated slab by slab."""
    from . import backend as B
    return B.partition_rcb(mesh.x[mesh.ien].mean(axis=1), nparts)


def element_partition_metis(mesh: M.Mesh, nparts: int) -> np.ndarray:
    """part[e] by the reference's own criterion (distribute.cpp:1683-1706 -> split_ -> ParMETIS_V3_PartMeshKway with
    ncommonnodes = eNoNb): METIS k-way partition of the dual graph, elements adjacent across a shared face."""
    from . import backend as B
    eNoNb = {4: 3, 8: 4, 10: 6}[mesh.ien.shape[1]]            # TET4 -> TRI3, HEX8 -> QUD4, TET10 -> TRI6 (consts.cpp element_type_to_elem_nonb)
    return B.partition_metis(mesh.ien, mesh.nNo, nparts, eNoNb)[0]


def element_partition(mesh: M.Mesh, nparts: int) -> np.ndarray:
    """part[e] for the generator's element order (6 tets per hex, hexes x-fastest then y then z)."""
    nx, ny, nz = mesh.shape
    per_layer = 6 * nx * ny
    part = np.empty(mesh.nEl, np.int32)
    for p, (k0, k1) in enumerate(slab_ranges(nz, nparts)):
        part[k0 * per_layer:k1 * per_layer] = p
    return part


# ------------------------------------------------------------------------------------------------
# per-rank pieces of a global case (parity tests: global case -> parts)
# ------------------------------------------------------------------------------------------------
def split_case(case, nparts: int, part: np.ndarray | None = None):
    """Cut a single-rank case (svfsiplus_b200.problem.pipe_case) into per-rank cases.

    Local node numbering = ascending global id (the order in which a rank meets its nodes is irrelevant
    to FSILS, which renumbers them itself).  Face vectors are restricted from the GLOBAL face vector,
    i.e. they are already overlap-summed like fsils_bc_create leaves them (bc.cpp:100-129).
    """
    m = case["mesh"]
    if part is None:
        part = element_partition(m, nparts)
    out = []
    for p in range(nparts):
        ien_g = m.ien[part == p]
        gN = np.unique(ien_g)
        g2l = np.full(m.nNo, -1, np.int64)
        g2l[gN] = np.arange(len(gN))
        ien = g2l[ien_g].astype(np.int32)
        rowPtr, colPtr = M.csr_pattern(ien, len(gN))
        faces = []
        for f in case["faces"]:
            loc = g2l[f["nodes"]]
            keep = loc >= 0
            faces.append(dict(name=f["name"], nodes=loc[keep].astype(np.int32), gnodes=np.asarray(f["nodes"])[keep],
                              dof=f["dof"], bGrp=f["bGrp"], val=np.ascontiguousarray(f["val"][keep])))
        lm = M.Mesh(x=np.ascontiguousarray(m.x[gN]), ien=np.ascontiguousarray(ien), faces={}, shape=m.shape)
        out.append(dict(mesh=lm, gNodes=gN.astype(np.int32), gnNo=m.nNo, rowPtr=rowPtr, colPtr=colPtr,
                        Ag=np.ascontiguousarray(case["Ag"][gN]), Yg=np.ascontiguousarray(case["Yg"][gN]),
                        Bf=np.ascontiguousarray(case["Bf"][gN]), props=case["props"], faces=faces,
                        res=case["res"], incL=case["incL"], rank=p, nranks=nparts))
    return out


# ------------------------------------------------------------------------------------------------
# fsils_lhs_create: node reordering + overlap lists
# ------------------------------------------------------------------------------------------------
def lhs_layout(rank: int, all_gnodes: list[np.ndarray], gnNo: int):
    """fsils_lhs_create's node ordering and overlap lists for `rank`: the product's native implementation behind the C ABI
    (b200_lhs_layout_*, csrc/lhs_layout.hpp).  `lhs_layout_numpy` below is the independent numpy restatement the tests
    compare it with (both are pinned against the compiled reference)."""
    from . import backend as B
    return B.lhs_layout(rank, all_gnodes, gnNo)


def lhs_layout_numpy(rank: int, all_gnodes: list[np.ndarray], gnNo: int):
    """The part of fsils_lhs_create (lhs.cpp:57-376) that is not a plain copy: for rank `rank`, given
    every rank's global node list in local order (the MPI_Allgatherv at lhs.cpp:156), return

      map[a]   local assembly id -> solver id                       (lhs.map)
      mynNo    rows this rank counts in dot products               (lhs.mynNo)
      shnNo    nodes shared with lower ranks                       (lhs.shnNo)
      reqs     [(peer, ptr)] with ptr = solver ids, ordered as the HIGHER rank of the pair lists them
               in its renumbered node order                        (lhs.cS[i].iP / .ptr)

    Single-rank: identity map, mynNo = nNo, no requests (lhs.cpp:96-121).
    """
    nT = len(all_gnodes)
    gN = np.asarray(all_gnodes[rank], np.int64)
    nNo = len(gN)
    if nT == 1:
        return dict(map=np.arange(nNo, dtype=np.int32), mynNo=nNo, shnNo=0, reqs=[])

    def ltg_of(r):
        """Renumbered (solver-order) global node list of rank r: the loop at lhs.cpp:171-224."""
        g = np.asarray(all_gnodes[r], np.int64)
        n = len(g)
        gtl = np.full(gnNo, -1, np.int64)
        gtl[g] = np.arange(n)
        taken = np.zeros(n, bool)
        low, high = [], []
        for i in range(nT - 1, -1, -1):              # ranks from the highest down, skipping r
            if i == r:
                continue
            ai = gtl[np.asarray(all_gnodes[i], np.int64)]
            ai = ai[ai >= 0]                          # local ids of the nodes rank i also holds, in rank i's order
            new = ai[~taken[ai]]
            taken[new] = True
            (low if i < r else high).append(g[new])
        low = np.concatenate(low) if low else np.zeros(0, np.int64)
        # nodes shared with higher ranks are written from the END backwards (lhs.cpp:198-199)
        high = np.concatenate(high)[::-1] if high else np.zeros(0, np.int64)
        interior = g[~taken]                          # remaining local nodes keep their relative order
        return np.concatenate([low, interior, high]), len(low), n - len(high)

    ltg, shnNo, mynNo = ltg_of(rank)
    gtl_new = np.full(gnNo, -1, np.int64)
    gtl_new[ltg] = np.arange(nNo)
    mp = gtl_new[gN].astype(np.int32)

    reqs = []
    for i in range(nT):
        if i == rank:
            continue
        # the list is built by the rank with the LOWER id from the HIGHER rank's renumbered order
        # (the `else` branch at lhs.cpp:336-360 runs on the lower rank, walks aNodes(:,iP) of the higher
        # one and sends the global ids; the higher rank just receives and maps them)
        gi = np.asarray(all_gnodes[i], np.int64)
        if not (gtl_new[gi] >= 0).any():
            continue                                  # no common node: no request (lhs.cpp:300-302)
        if i < rank:                                  # this rank is the higher one: its own order
            present = np.zeros(gnNo, bool)
            present[gi] = True
            shared = ltg[present[ltg]]
        else:                                         # walk the higher rank's renumbered list
            ltg_hi = ltg_of(i)[0]
            shared = ltg_hi[gtl_new[ltg_hi] >= 0]
        reqs.append((i, gtl_new[shared].astype(np.int32)))
    return dict(map=mp, mynNo=int(mynNo), shnNo=int(shnNo), reqs=reqs)


# ------------------------------------------------------------------------------------------------
# backend set-up of one rank
# ------------------------------------------------------------------------------------------------
def setup_rank_backend(part, layout, device, uid=None, be=None):
    """What initialize() + fsils_lhs_create + fsils_bc_create + add_eq_linear_algebra do on one rank."""
    from . import backend as B
    if be is None:
        be = B.Backend(device)
    if part["nranks"] > 1:
        be.comm_init(part["rank"], part["nranks"], uid)
    m = part["mesh"]
    be.lhs_create(part["gnNo"], part["rowPtr"], part["colPtr"], map=layout["map"], mynNo=layout["mynNo"],
                  reqs=layout["reqs"], nFaces=len(part["faces"]))
    shared = part.get("face_shared")
    for i, f in enumerate(part["faces"]):
        sh = bool(shared[i]) if shared is not None else False
        be.face_set(i, layout["map"][f["nodes"]] if len(f["nodes"]) else np.zeros(0, np.int32), f["dof"], f["bGrp"],
                    f["val"], shared=sh)
    be.mesh_set(m.ien, m.x)
    return be


def face_shared_flags(parts):
    """lhs.face[].sharedFlag: a face is shared when more than one rank holds nodes of it (bc.cpp:96-103)."""
    nF = len(parts[0]["faces"])
    return [sum(1 for p in parts if len(p["faces"][i]["nodes"]) > 0) > 1 for i in range(nF)]


# ------------------------------------------------------------------------------------------------
# weak-scaling workload generated rank-locally (bench.py --gpus N)
# ------------------------------------------------------------------------------------------------
def weak_dims(base_dims, world):
    """N GPUs: the SAME pipe refined to N times the tets (BASELINE.json configs[2]: "same pipe refined to ~80M
    tets" on 8 GPUs), every direction scaled by N^(1/3): N = 8 gives 192 x 192 x 362 = 80 068 608 TET4 exactly,
    N = 2 -> 121 x 121 x 228 (20.0 M), N = 4 -> 152 x 152 x 287 (39.8 M).  Cut into z-slabs of whole layers."""
    nx, ny, nz = base_dims
    f = float(world) ** (1.0 / 3.0)
    return (int(round(nx * f)), int(round(ny * f)), int(round(nz * f)))


def generation_blocks(world: int) -> int:
    """Number of independently generated z-blocks the synthetic pipe is composed of.  It does NOT depend on how many ranks
    own them when the rank count divides 8, so the global mesh and state are bit-identical for 1, 2, 4 and 8 GPUs
    (strong-scaling runs and the single-GPU line solve the same system)."""
    return 8 if 8 % world == 0 else world


def _pipe_block(dims, b, blocks, radius, length):
    """Generation block b of the global pipe: its hex layers, jittered on its own with a block-dependent seed; the two
    interface planes are left unjittered and carry state from a plane-keyed stream, so neighbouring blocks agree bit for bit."""
    nx, ny, nz = dims
    k0, k1 = slab_ranges(nz, blocks)[b]
    plane = (nx + 1) * (ny + 1)
    dz = length / nz
    m = M.pipe_mesh(nx, ny, k1 - k0, radius=radius, length=(k1 - k0) * dz, jitter=0.1, seed=1234 + b)
    m.x[:, 2] += k0 * dz
    m.x[:plane, 2] = k0 * dz                      # the same expression on both owners of an interface plane
    m.x[m.nNo - plane:, 2] = k1 * dz
    Ag, Yg, Bf = M.pipe_state(m, radius=radius, length=length, seed_y=2024 + b, seed_a=2025 + b)
    for kk, sl in ((k0, slice(0, plane)), (k1, slice(m.nNo - plane, m.nNo))):
        if 0 < kk < nz:
            rng = np.random.default_rng(777000 + kk)
            r2 = (m.x[sl, 0] ** 2 + m.x[sl, 1] ** 2) / radius ** 2
            Yg[sl, :3] = 0.0
            Yg[sl, 2] = 20.0 * np.clip(1.0 - r2, 0.0, None)
            Yg[sl, :3] += 0.2 * rng.standard_normal((plane, 3))
            Yg[sl, 3] = 100.0 * (1.0 - m.x[sl, 2] / length) + rng.standard_normal(plane)
            Ag[sl, :4] = 10.0 * rng.standard_normal((plane, 4))
    return m, Ag, Yg, Bf, k0, k1


def local_slab_case(dims, rank, world, *, radius=1.0, length=10.0, pattern=None, blocks=None):
    """The rank's z-slab of the global pipe `dims`, generated without ever building the global mesh.

    Node (i,j,k) has global id k*(nx+1)*(ny+1) + j*(nx+1) + i.  The pipe is composed of `blocks` generation blocks
    (generation_blocks: 8 for 1, 2, 4, 8 ranks) of whole hex layers; a rank owns blocks/world consecutive ones and merges
    them (shared interface planes are identical by construction).  world = 1 therefore yields the SAME global mesh and
    state that 2, 4 or 8 ranks hold in pieces.
    Returns a per-rank case like split_case's plus `all_gnodes` (analytic, no communication needed).
    pattern: optional callable (nNo, ien) -> (rowPtr, colPtr), e.g. the device-side lhsa (Backend.pattern); the host
    construction (numpy) takes ~30 s at 10 M tets.
    """
    from . import backend as B
    nx, ny, nz = dims
    if blocks is None:
        blocks = generation_blocks(world)
    if blocks % world != 0 or blocks > nz:
        raise ValueError("blocks must be a multiple of the rank count and at most nz")
    per = blocks // world
    plane = (nx + 1) * (ny + 1)
    xs, iens, As, Ys, Bs = [], [], [], [], []
    off = 0
    K0 = K1 = None
    out_tris = None
    for b in range(rank * per, (rank + 1) * per):
        mb, Ag_b, Yg_b, Bf_b, k0, k1 = _pipe_block(dims, b, blocks, radius, length)
        first = K0 is None
        if first:
            K0 = k0
        K1 = k1
        skip = 0 if first else plane                  # the shared plane comes from the lower block
        xs.append(mb.x[skip:]); As.append(Ag_b[skip:]); Ys.append(Yg_b[skip:]); Bs.append(Bf_b[skip:])
        iens.append(mb.ien + np.int32(off - skip))
        out_tris = mb.faces["outlet"]["tris"] + np.int32(off - skip)
        off += mb.nNo - skip
    x = np.ascontiguousarray(np.concatenate(xs)); ien = np.ascontiguousarray(np.concatenate(iens).astype(np.int32))
    Ag = np.ascontiguousarray(np.concatenate(As)); Yg = np.ascontiguousarray(np.concatenate(Ys)); Bf = np.ascontiguousarray(np.concatenate(Bs))
    del xs, iens, As, Ys, Bs
    nNo = x.shape[0]
    assert nNo == plane * (K1 - K0 + 1)
    nid = np.arange(nNo, dtype=np.int32)
    I = nid % (nx + 1); J = (nid // (nx + 1)) % (ny + 1)
    wall = nid[(I == 0) | (I == nx) | (J == 0) | (J == ny)]
    m = M.Mesh(x=x, ien=ien, faces={}, shape=(nx, ny, nz))
    gN = (np.arange(nNo, dtype=np.int64) + K0 * plane).astype(np.int32)
    rowPtr, colPtr = pattern(nNo, ien) if pattern else M.csr_pattern(ien, nNo)
    am, af, gam = M.gen_alpha(0.5)
    dt = 0.005
    props = dict(dt=dt, am=am, af=af, gam=gam, rho=1.06, mu=0.04)
    inlet = nid[:plane] if K0 == 0 else np.zeros(0, np.int32)
    outlet = nid[nNo - plane:] if K1 == nz else np.zeros(0, np.int32)
    out_val = (M.face_normal_integral(x, out_tris, outlet) if K1 == nz else np.zeros((0, 3)))
    faces = [dict(name="lumen_inlet", nodes=inlet, dof=3, bGrp=B.BC_DIR, val=np.zeros((len(inlet), 3))),
             dict(name="lumen_wall", nodes=wall, dof=3, bGrp=B.BC_DIR, val=np.zeros((len(wall), 3))),
             dict(name="lumen_outlet", nodes=outlet, dof=3, bGrp=B.BC_NEU, val=out_val)]
    res = np.array([0.0, 0.0, gam * dt * (121.0 + 1212.0)])
    br = slab_ranges(nz, blocks)
    all_gnodes = [np.arange(br[r * per][0] * plane, (br[(r + 1) * per - 1][1] + 1) * plane, dtype=np.int32) for r in range(world)]
    part = dict(mesh=m, gNodes=gN, gnNo=plane * (nz + 1), rowPtr=rowPtr, colPtr=colPtr, Ag=Ag, Yg=Yg, Bf=Bf,
                props=props, faces=faces, res=res, incL=np.array([1, 1, 1], np.int32), rank=rank, nranks=world,
                face_shared=[False, world > 1, False])
    return part, all_gnodes


def setup_distributed_case(dims, rank, world, local_device, dist=None):
    """bench.py entry: rank-local slab, FSILS layout, NCCL communicator bootstrapped over torch.distributed."""
    import torch
    from . import backend as B
    be = B.Backend(local_device)
    part, all_gnodes = local_slab_case(dims, rank, world, pattern=lambda n, ien: be.pattern(n, [ien]))     # lhsa on the device
    layout = lhs_layout(rank, all_gnodes, part["gnNo"])
    uid = None
    if world > 1:
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.from_numpy(B.unique_id().copy())
        t = t.cuda() if dist.get_backend() == "nccl" else t
        dist.broadcast(t, src=0)
        uid = t.cpu().numpy()
    be = setup_rank_backend(part, layout, local_device, uid, be=be)
    return part, be
