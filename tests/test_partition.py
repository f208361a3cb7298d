"""Host logic of the multi-GPU path on CPU: the slab partition, the numpy restatement of
fsils_lhs_create's node reordering and overlap lists against the COMPILED REFERENCE run with
threads-as-ranks (oracle/_ref), and the same set-up under torch.distributed (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import needs_ref
from util import ROOT

from svfsiplus_b200 import partition as PT
from svfsiplus_b200 import problem as P


@pytest.mark.parametrize("nparts", [2, 3, 4])
def test_split_covers_mesh(nparts):
    case = P.pipe_case(4, 4, 9)
    parts = PT.split_case(case, nparts)
    m = case["mesh"]
    assert sum(p["mesh"].nEl for p in parts) == m.nEl
    seen = np.zeros(m.nNo, int)
    for p in parts:
        seen[p["gNodes"]] += 1
        assert (p["mesh"].x == m.x[p["gNodes"]]).all()
        # local connectivity points at the same global nodes
        assert (np.sort(np.unique(p["gNodes"][p["mesh"].ien])) == np.sort(p["gNodes"])).all()
    assert (seen >= 1).all() and (seen <= 2).all()       # slabs: interface planes are held twice


@needs_ref
@pytest.mark.parametrize("nparts", [2, 3, 4])
def test_lhs_layout_equals_reference_fsils_lhs_create(nparts):
    from oracle import ref
    case = P.pipe_case(4, 4, 9)
    parts = PT.split_case(case, nparts)
    rr = ref.RefRanks([dict(gnNo=p["gnNo"], gNodes=p["gNodes"], rowPtr=p["rowPtr"], colPtr=p["colPtr"], faces=[])
                       for p in parts])
    allg = [p["gNodes"] for p in parts]
    for r in range(nparts):
        info = rr.info(r)
        for impl in (PT.lhs_layout, PT.lhs_layout_numpy):        # native (C ABI) and the numpy restatement
            lay = impl(r, allg, parts[r]["gnNo"])
            assert lay["mynNo"] == info["mynNo"] and lay["shnNo"] == info["shnNo"]
            assert (lay["map"] == info["map"]).all()
            assert len(lay["reqs"]) == info["nReq"]
            for (pa, ptra), (pb, ptrb) in zip(lay["reqs"], info["reqs"]):
                assert pa == pb and (ptra == ptrb).all()
    rr.close()


@needs_ref
def test_lhs_layout_equals_reference_irregular_partition():
    """A partition that is NOT slabs (round-robin blocks of elements): more than two owners per node."""
    from oracle import ref
    case = P.pipe_case(3, 3, 5)
    m = case["mesh"]
    part = (np.arange(m.nEl) // 7) % 3
    parts = PT.split_case(case, 3, part.astype(np.int32))
    rr = ref.RefRanks([dict(gnNo=p["gnNo"], gNodes=p["gNodes"], rowPtr=p["rowPtr"], colPtr=p["colPtr"], faces=[])
                       for p in parts])
    allg = [p["gNodes"] for p in parts]
    for r in range(3):
        info = rr.info(r)
        for impl in (PT.lhs_layout, PT.lhs_layout_numpy):
            lay = impl(r, allg, parts[r]["gnNo"])
            assert lay["mynNo"] == info["mynNo"] and lay["shnNo"] == info["shnNo"]
            assert (lay["map"] == info["map"]).all()
            assert [q[0] for q in lay["reqs"]] == [q[0] for q in info["reqs"]]
            for (pa, ptra), (pb, ptrb) in zip(lay["reqs"], info["reqs"]):
                assert (ptra == ptrb).all()
    rr.close()


@pytest.mark.parametrize("nparts,seed", [(1, 0), (2, 1), (5, 2), (8, 3)])
def test_native_layout_equals_numpy_restatement_random_partitions(nparts, seed):
    """Random element -> rank maps (many owners per node, ranks without common nodes, local orders shuffled): the native
    implementation behind the C ABI and the numpy restatement agree integer for integer; the map is a permutation and the
    overlap lists of a pair name the same global nodes in the same order on both sides."""
    rng = np.random.default_rng(seed)
    case = P.pipe_case(4, 4, 8)
    m = case["mesh"]
    # blocks of consecutive elements so that distant ranks share nothing
    part = np.minimum((np.arange(m.nEl) * nparts) // m.nEl + rng.integers(0, 2, m.nEl), nparts - 1).astype(np.int32)
    allg = []
    for r in range(nparts):
        g = np.unique(m.ien[part == r]).astype(np.int32)
        rng.shuffle(g)                                        # local order is arbitrary
        allg.append(g)
    lays = []
    for r in range(nparts):
        a = PT.lhs_layout(r, allg, m.nNo)
        b = PT.lhs_layout_numpy(r, allg, m.nNo)
        assert a["mynNo"] == b["mynNo"] and a["shnNo"] == b["shnNo"] and (a["map"] == b["map"]).all()
        assert [q[0] for q in a["reqs"]] == [q[0] for q in b["reqs"]]
        for (_, pa), (_, pb) in zip(a["reqs"], b["reqs"]):
            assert (pa == pb).all()
        assert sorted(a["map"].tolist()) == list(range(len(allg[r])))
        lays.append(a)
    # every node is counted by exactly one rank's [0, mynNo)
    owners = np.zeros(m.nNo, int)
    for r, lay in enumerate(lays):
        inv = np.empty(len(allg[r]), np.int64)
        inv[lay["map"]] = allg[r]
        owners[inv[:lay["mynNo"]]] += 1
        for peer, ptr in lay["reqs"]:
            back = dict(lays[peer]["reqs"])[r]
            inv_p = np.empty(len(allg[peer]), np.int64)
            inv_p[lays[peer]["map"]] = allg[peer]
            assert (inv[ptr] == inv_p[back]).all()
    held = np.zeros(m.nNo, bool)
    for g in allg:
        held[g] = True
    assert (owners[held] == 1).all()


def test_native_layout_reports_bad_input():
    with pytest.raises(RuntimeError, match="outside"):
        PT.lhs_layout(0, [np.array([0, 1, 7], np.int32), np.array([1, 2], np.int32)], 5)
    with pytest.raises(RuntimeError, match="twice"):
        PT.lhs_layout(0, [np.array([0, 1, 1], np.int32), np.array([1, 2], np.int32)], 5)


def test_local_slab_case_matches_neighbours():
    """Rank-local generation (bench --gpus N): interface planes agree between neighbours."""
    dims = (4, 4, 8)
    a, ga = PT.local_slab_case(dims, 0, 2)
    b, gb = PT.local_slab_case(dims, 1, 2)
    plane = 25
    assert (a["gNodes"][-plane:] == b["gNodes"][:plane]).all()
    assert np.array_equal(a["mesh"].x[-plane:], b["mesh"].x[:plane])
    assert np.array_equal(a["Yg"][-plane:], b["Yg"][:plane]) and np.array_equal(a["Ag"][-plane:], b["Ag"][:plane])
    la = PT.lhs_layout(0, ga, a["gnNo"]); lb = PT.lhs_layout(1, gb, b["gnNo"])
    assert la["mynNo"] == a["mesh"].nNo - plane and lb["shnNo"] == plane and lb["mynNo"] == b["mesh"].nNo
    # the two ends of a request list name the same global nodes in the same order
    inv_a = np.empty(a["mesh"].nNo, int); inv_a[la["map"]] = np.arange(a["mesh"].nNo)
    inv_b = np.empty(b["mesh"].nNo, int); inv_b[lb["map"]] = np.arange(b["mesh"].nNo)
    assert (a["gNodes"][inv_a[la["reqs"][0][1]]] == b["gNodes"][inv_b[lb["reqs"][0][1]]]).all()


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from svfsiplus_b200 import partition as PT, problem as P
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
case = P.pipe_case(4, 4, 9)
part = PT.split_case(case, world)[rank]
# what a multi-process run does: every rank contributes its node list, all ranks derive their layout
lst = [None] * world
dist.all_gather_object(lst, part["gNodes"])
lay = PT.lhs_layout(rank, lst, part["gnNo"])
# consistency across ranks: each pair of request lists has equal length and names the same global nodes
inv = np.empty(len(part["gNodes"]), int); inv[lay["map"]] = np.arange(len(part["gNodes"]))
mine = {peer: part["gNodes"][inv[ptr]] for peer, ptr in lay["reqs"]}
allm = [None] * world
dist.all_gather_object(allm, mine)
for peer, g in mine.items():
    assert (allm[peer][rank] == g).all()
owned = torch.tensor([lay["mynNo"]], dtype=torch.int64)
dist.all_reduce(owned)
assert int(owned) == case["mesh"].nNo, (int(owned), case["mesh"].nNo)     # every global node is counted once
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_layout_under_gloo_world_size_2(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(w), ROOT],
                       capture_output=True, text=True, timeout=600, env=env)
    assert "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nparts", [1, 2, 3, 5, 8])
def test_rcb_partition_is_balanced_deterministic_and_cuts_the_pipe_across_its_axis(nparts):
    case = P.pipe_case(4, 4, 12)
    m = case["mesh"]
    part = PT.element_partition_rcb(m, nparts)
    assert part.min() == 0 and part.max() == nparts - 1
    counts = np.bincount(part, minlength=nparts)
    assert counts.max() - counts.min() <= 1 and counts.sum() == m.nEl
    assert np.array_equal(part, PT.element_partition_rcb(m, nparts))
    # the pipe is ten times longer than wide: every cut is across z, parts are ordered along the axis
    zc = m.x[m.ien].mean(axis=1)[:, 2]
    for p in range(nparts - 1):
        assert zc[part == p].max() <= zc[part == p + 1].min()


@needs_ref
def test_rcb_partition_layout_equals_reference_fsils_lhs_create():
    """An RCB partition of a block (cuts in several directions, nodes with up to four owners) through split_case and the native
    layout against the reference's fsils_lhs_create on threads-as-ranks."""
    from oracle import ref
    from svfsiplus_b200 import mesh as M
    case = P.pipe_case(5, 5, 5)
    m = case["mesh"]
    # make the domain cube-like so that RCB cuts along different axes
    m2 = M.Mesh(x=m.x * np.array([1.0, 1.0, 0.2]), ien=m.ien, faces=m.faces, shape=m.shape)
    part = PT.element_partition_rcb(m2, 4)
    zc, xc = m2.x[m2.ien].mean(axis=1)[:, 2], m2.x[m2.ien].mean(axis=1)[:, 0]
    assert len({(zc[part == p].mean() > zc.mean(), xc[part == p].mean() > xc.mean()) for p in range(4)}) >= 3      # not slabs
    parts = PT.split_case(case, 4, part)
    rr = ref.RefRanks([dict(gnNo=p["gnNo"], gNodes=p["gNodes"], rowPtr=p["rowPtr"], colPtr=p["colPtr"], faces=[]) for p in parts])
    allg = [p["gNodes"] for p in parts]
    for r in range(4):
        info = rr.info(r)
        lay = PT.lhs_layout(r, allg, parts[r]["gnNo"])
        assert lay["mynNo"] == info["mynNo"] and lay["shnNo"] == info["shnNo"] and (lay["map"] == info["map"]).all()
        assert [q[0] for q in lay["reqs"]] == [q[0] for q in info["reqs"]]
        for (_, pa), (_, pb) in zip(lay["reqs"], info["reqs"]):
            assert (pa == pb).all()
    rr.close()


@needs_ref
@pytest.mark.parametrize("nparts,seed", [(2, 11), (5, 12), (6, 13)])
def test_native_layout_equals_reference_on_random_partitions(nparts, seed):
    """Random element -> rank maps (ragged part sizes, nodes with many owners, possibly ranks that share nothing) through split_case
    and the native layout, integer for integer against fsils_lhs_create on threads-as-ranks."""
    from oracle import ref
    rng = np.random.default_rng(seed)
    case = P.pipe_case(3, 3, 7)
    m = case["mesh"]
    part = rng.integers(0, nparts, m.nEl).astype(np.int32)
    part[:nparts] = np.arange(nparts)                          # no empty rank
    parts = PT.split_case(case, nparts, part)
    rr = ref.RefRanks([dict(gnNo=p["gnNo"], gNodes=p["gNodes"], rowPtr=p["rowPtr"], colPtr=p["colPtr"], faces=[]) for p in parts])
    allg = [p["gNodes"] for p in parts]
    for r in range(nparts):
        info = rr.info(r)
        lay = PT.lhs_layout(r, allg, parts[r]["gnNo"])
        assert lay["mynNo"] == info["mynNo"] and lay["shnNo"] == info["shnNo"] and (lay["map"] == info["map"]).all()
        assert [q[0] for q in lay["reqs"]] == [q[0] for q in info["reqs"]]
        for (_, pa), (_, pb) in zip(lay["reqs"], info["reqs"]):
            assert (pa == pb).all()
    rr.close()


# ---------------------------------------------------------------------------------------------------
# the reference's own partition criterion (METIS dual graph) and the block-composed benchmark pipe
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nparts", [2, 3, 4, 8])
def test_metis_dual_graph_partition_is_valid_and_balanced(nparts):
    """b200_partition_metis (distribute.cpp:1683-1706: ncommonnodes = eNoNb): every element gets a part, parts are balanced
    within METIS' default 3 % tolerance (+ slack for tiny meshes), the edge cut is far below a random assignment's, and the
    FSILS layout (b200_lhs_layout_*) of the resulting irregular partition is consistent between neighbours."""
    from svfsiplus_b200 import backend as B
    from svfsiplus_b200 import mesh as M
    m = M.pipe_mesh(12, 12, 24)
    part, cut = B.partition_metis(m.ien, m.nNo, nparts, 3)
    assert part.min() == 0 and part.max() == nparts - 1
    cnt = np.bincount(part, minlength=nparts)
    assert cnt.max() <= 1.06 * cnt.mean()
    # faces between parts, counted independently: tets sharing 3 nodes
    from collections import defaultdict
    faces = defaultdict(list)
    for e, el in enumerate(m.ien):
        for tri in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)):
            faces[tuple(sorted(el[list(tri)]))].append(e)
    mycut = sum(1 for v in faces.values() if len(v) == 2 and part[v[0]] != part[v[1]])
    assert mycut == cut
    inner = sum(1 for v in faces.values() if len(v) == 2)
    assert cut < 0.25 * inner * (1.0 - 1.0 / nparts)
    case = dict(mesh=m, Ag=np.zeros((m.nNo, 4)), Yg=np.zeros((m.nNo, 4)), Bf=np.zeros((m.nNo, 3)), props={}, faces=[], res=None, incL=None)
    parts = PT.split_case(case, nparts, part)
    lays = [PT.lhs_layout(r, [p["gNodes"] for p in parts], m.nNo) for r in range(nparts)]
    for r, lay in enumerate(lays):
        for peer, lst in lay["reqs"]:
            back = dict(lays[peer]["reqs"])[r]
            inv_r = np.empty(len(lay["map"]), np.int64); inv_r[lay["map"]] = np.arange(len(lay["map"]))
            inv_p = np.empty(len(lays[peer]["map"]), np.int64); inv_p[lays[peer]["map"]] = np.arange(len(lays[peer]["map"]))
            # the i-th entry of both lists is the same global node (fsils_commuv pairs them by position)
            assert np.array_equal(parts[r]["gNodes"][inv_r[lst]], parts[peer]["gNodes"][inv_p[back]])


def test_metis_partition_errors():
    from svfsiplus_b200 import backend as B
    with pytest.raises(RuntimeError, match="out of range"):
        B.partition_metis(np.array([[0, 1, 2, 9]], np.int32), 4, 2, 3)


def test_block_composed_pipe_is_identical_for_every_rank_count():
    """bench.py's workload (partition.local_slab_case): the pipe is composed of 8 generation blocks, so the pieces 2, 4 and 8 ranks
    hold are bit-identical restrictions of the mesh and state one rank holds - strong-scaling runs solve the same system."""
    dims = (6, 6, 20)
    p1, g1 = PT.local_slab_case(dims, 0, 1)
    assert len(g1) == 1 and p1["mesh"].nNo == 7 * 7 * 21 and p1["mesh"].nEl == 6 * 6 * 6 * 20
    from svfsiplus_b200 import mesh as M
    assert (M.tet_volumes(p1["mesh"].x, p1["mesh"].ien) > 0).all()
    for w in (2, 4, 8):
        els = 0
        for r in range(w):
            pr, ag = PT.local_slab_case(dims, r, w)
            g = pr["gNodes"]
            assert np.array_equal(np.asarray(ag[r]), g)
            for k in ("Ag", "Yg", "Bf"):
                assert np.array_equal(pr[k], p1[k][g])
            assert np.array_equal(pr["mesh"].x, p1["mesh"].x[g])
            gi = g[pr["mesh"].ien]
            assert np.array_equal(gi, p1["mesh"].ien[els:els + len(gi)])
            els += len(gi)
            for f, f1 in zip(pr["faces"], p1["faces"]):
                assert set(g[f["nodes"]]) <= set(f1["nodes"])
            if len(pr["faces"][2]["nodes"]):
                assert np.array_equal(pr["faces"][2]["val"], p1["faces"][2]["val"])
        assert els == p1["mesh"].nEl


# ---------------------------------------------------------------------------------------------------
# host-side tables of the device transport and of the row-tile kernels (csrc/lhs_layout.hpp, through tests/hostlogic)
# ---------------------------------------------------------------------------------------------------
def _hl():
    from util import hostlogic
    return hostlogic()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_halo_source_lists_keep_request_order(seed):
    """fsils_commuv adds the received values request by request (in_commu.cpp:150-168); the device adds per overlap row, so every
    row's sources must be listed in request order, each (request, position) exactly once."""
    import ctypes as C
    rng = np.random.default_rng(seed)
    nNo, nReq = 200, 4
    lists = [rng.choice(nNo, size=rng.integers(5, 60), replace=False).astype(np.int32) for _ in range(nReq)]
    lists[2] = np.concatenate([lists[2], lists[0][:7]]).astype(np.int32)           # rows shared by two and three requests
    lists[3] = np.unique(np.concatenate([lists[3], lists[0][:4], lists[1][:5]])).astype(np.int32)
    lists[2] = np.array(list(dict.fromkeys(lists[2].tolist())), np.int32)
    req_n = np.array([len(l) for l in lists], np.int32)
    cat = np.concatenate(lists).astype(np.int32)
    tot = int(req_n.sum())
    nh = C.c_int(0)
    node = np.zeros(nNo + 1, np.int32); ptr = np.zeros(nNo + 2, np.int32); sr = np.zeros(tot, np.int32); sp = np.zeros(tot, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert _hl().hl_halo_sources(nNo, nReq, p(req_n), p(cat), C.byref(nh), p(node), p(ptr), p(sr), p(sp)) == 0
    n = nh.value
    assert list(node[:n]) == sorted(set(cat.tolist())) and ptr[n] == tot
    seen = set()
    for k in range(n):
        reqs = sr[ptr[k]:ptr[k+1]]
        assert list(reqs) == sorted(reqs) and len(set(reqs.tolist())) == len(reqs)      # request order, one source per request
        for e in range(ptr[k], ptr[k+1]):
            assert lists[sr[e]][sp[e]] == node[k]
            seen.add((int(sr[e]), int(sp[e])))
    assert len(seen) == tot
    # emulate the exchange: every overlap row gets the sum of its sources added in request order = what the request loop does
    V = rng.standard_normal(nNo)
    rbuf = [rng.standard_normal(len(l)) for l in lists]
    ref = V.copy()
    for i, l in enumerate(lists):
        for j, r in enumerate(l):
            ref[r] += rbuf[i][j]
    got = V.copy()
    for k in range(n):
        s = got[node[k]]
        for e in range(ptr[k], ptr[k+1]):
            s += rbuf[sr[e]][sp[e]]
        got[node[k]] = s
    assert np.array_equal(got, ref)                                                    # bit for bit: same order of additions
    bad = np.array([nNo], np.int32)
    assert _hl().hl_halo_sources(nNo, 1, p(np.array([1], np.int32)), p(bad), C.byref(nh), p(node), p(ptr), p(sr), p(sp)) == 1


def test_row_tiles_cover_the_rows_and_respect_caps_and_cuts():
    import ctypes as C
    from svfsiplus_b200 import mesh as M
    m = M.pipe_mesh(8, 8, 16)
    rowPtr, _ = M.csr_pattern(m.ien, m.nNo)
    rowPtr = np.ascontiguousarray(rowPtr, np.int32)
    nNo = m.nNo
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for ovA, ovB, max_rows, cap in ((0, nNo, 128, 1152), (81, nNo - 81, 128, 1152), (81, nNo - 81, 16, 200), (0, nNo, 128, 64)):
        tr = np.zeros(nNo + 1, np.int32); at = np.zeros(4, np.int32)
        nt = _hl().hl_row_tiles(nNo, p(rowPtr), ovA, ovB, max_rows, cap, p(tr), p(at))
        assert nt > 0
        t = tr[:nt + 1]
        assert t[0] == 0 and t[-1] == nNo and (np.diff(t) > 0).all() and (np.diff(t) <= max_rows).all()
        for a, b in zip(t[:-1], t[1:]):
            assert rowPtr[b] - (rowPtr[a] & ~3) <= cap - 4
            assert not (a < ovA < b) and not (a < ovB < b)                             # tiles never straddle the segment cuts
        assert t[at[0]] == 0 and at[3] == nt
        if ovA > 0:
            assert t[at[1]] == ovA and t[at[2]] == ovB
    # a row that does not fit into a tile: no tiles (the per-lane kernels are used)
    tr = np.zeros(nNo + 1, np.int32); at = np.zeros(4, np.int32)
    assert _hl().hl_row_tiles(nNo, p(rowPtr), 0, nNo, 128, 12, p(tr), p(at)) == 0
