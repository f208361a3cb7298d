"""Synthetic meshes and states for the nonlinear-step hot path.

The reference reads its meshes through VTK (Code/Source/solver/vtk_xml.cpp, load_msh.cpp); the
shipped test meshes are Git-LFS stubs, so every workload here is generated: a structured block of
hexes, each split into 6 Kuhn tets (conforming), the square cross-section mapped onto a disc
(cylinder "pipe" of tests/cases/fluid/pipe_RCR_3d, CGS units).  Arrays follow the reference's
column-major layout seen from C: ``x[a, :]`` are the nsd coordinates of node ``a`` (= ``x(:,a)`` of
Array<double> x(nsd,tnNo), Code/Source/solver/ComMod.h:1563) and ``ien[e, :]`` the nodes of element
``e`` (= ``IEN(:,e)``, ComMod.h:893).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    x: np.ndarray                 # (nNo, 3) float64
    ien: np.ndarray               # (nEl, eNoN) int32
    faces: dict = field(default_factory=dict)   # name -> dict(nodes=int32[], tris=int32[nf,3] | None)
    shape: tuple = ()

    @property
    def nNo(self) -> int:
        return self.x.shape[0]

    @property
    def nEl(self) -> int:
        return self.ien.shape[0]


def _kuhn_tets(nx: int, ny: int, nz: int) -> np.ndarray:
    """6 Kuhn tets per hex around the (0,0,0)-(1,1,1) diagonal; translation invariant => conforming."""
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # element order: x fastest, then y, then z (same as node order)
    base = (i * sx + j * sy + k * sz).transpose(2, 1, 0).reshape(-1).astype(np.int64)
    strides = (sx, sy, sz)
    tets = []
    for perm in itertools.permutations(range(3)):
        o1 = strides[perm[0]]
        o2 = o1 + strides[perm[1]]
        o3 = sx + sy + sz
        tets.append(np.stack([base, base + o1, base + o2, base + o3], axis=1))
    ien = np.stack(tets, axis=1).reshape(-1, 4)      # the 6 tets of a hex are consecutive
    return ien


def _fix_orientation(x: np.ndarray, ien: np.ndarray) -> np.ndarray:
    """Swap two nodes of every tet with negative Jacobian so that det(dx/dxi) > 0 everywhere."""
    p0 = x[ien[:, 3]]
    a = x[ien[:, 0]] - p0
    b = x[ien[:, 1]] - p0
    c = x[ien[:, 2]] - p0
    det = np.einsum("ij,ij->i", a, np.cross(b, c))
    neg = det < 0
    ien = ien.copy()
    ien[neg, 0], ien[neg, 1] = ien[neg, 1].copy(), ien[neg, 0].copy()
    return ien


def tet_volumes(x: np.ndarray, ien: np.ndarray) -> np.ndarray:
    p0 = x[ien[:, 3]]
    a = x[ien[:, 0]] - p0
    b = x[ien[:, 1]] - p0
    c = x[ien[:, 2]] - p0
    return np.einsum("ij,ij->i", a, np.cross(b, c)) / 6.0


def node_hmin(x: np.ndarray, ien: np.ndarray) -> np.ndarray:
    """Local spacing per node: the shortest edge of any incident tet (TET4 connectivity)."""
    hmin = np.full(x.shape[0], np.inf)
    for a, b in itertools.combinations(range(4), 2):
        d = np.linalg.norm(x[ien[:, a]] - x[ien[:, b]], axis=1)
        np.minimum.at(hmin, ien[:, a], d)
        np.minimum.at(hmin, ien[:, b], d)
    return hmin


def pipe_mesh(nx: int, ny: int, nz: int, radius: float = 1.0, length: float = 10.0,
              jitter: float = 0.1, seed: int = 1234) -> Mesh:
    """Cylinder of TET4: nx*ny*nz hexes -> 6*nx*ny*nz tets, (nx+1)(ny+1)(nz+1) nodes.

    P10 of SURVEY.md §8d is pipe_mesh(96, 96, 181); P80 is pipe_mesh(192, 192, 362).
    Faces: 'inlet' (z=0), 'outlet' (z=L, with its boundary triangles) and 'wall'.
    """
    u = np.linspace(-1.0, 1.0, nx + 1)
    v = np.linspace(-1.0, 1.0, ny + 1)
    w = np.linspace(0.0, length, nz + 1)
    W, V, U = np.meshgrid(w, v, u, indexing="ij")          # node order: x fastest
    # concentric (Shirley-Chiu) square->disc map: radius = max(|u|,|v|), angle linear along the
    # square's perimeter.  Unlike the elliptical map its Jacobian stays bounded away from zero at the
    # four square corners, so the corner cells keep an O(h^2) area.
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.abs(U) >= np.abs(V)
        r = np.where(a, U, V)
        phi = np.where(a, (np.pi / 4.0) * np.where(U != 0.0, V / np.where(U == 0.0, 1.0, U), 0.0),
                       (np.pi / 2.0) - (np.pi / 4.0) * np.where(V != 0.0, U / np.where(V == 0.0, 1.0, V), 0.0))
    X = radius * r * np.cos(phi)
    Y = radius * r * np.sin(phi)
    x = np.stack([X.reshape(-1), Y.reshape(-1), W.reshape(-1)], axis=1)

    ii = np.arange(nx + 1)
    jj = np.arange(ny + 1)
    kk = np.arange(nz + 1)
    K, J, I = np.meshgrid(kk, jj, ii, indexing="ij")
    I = I.reshape(-1); J = J.reshape(-1); K = K.reshape(-1)
    on_wall = (I == 0) | (I == nx) | (J == 0) | (J == ny)
    on_in = K == 0
    on_out = K == nz
    interior = ~(on_wall | on_in | on_out)

    ien = _kuhn_tets(nx, ny, nz)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        # local spacing: shortest edge of any incident tet
        e = ien
        hmin = np.full(x.shape[0], np.inf)
        for a, b in itertools.combinations(range(4), 2):
            d = np.linalg.norm(x[e[:, a]] - x[e[:, b]], axis=1)
            np.minimum.at(hmin, e[:, a], d)
            np.minimum.at(hmin, e[:, b], d)
        dx = rng.uniform(-1.0, 1.0, size=x.shape) * (jitter * hmin)[:, None]
        dx[~interior] = 0.0
        # thin cells along the map's diagonals must not fold: halve the displacement of the nodes of
        # any element that loses more than 70% of its volume (or inverts) until none does
        v0 = tet_volumes(x, ien)
        for _ in range(12):
            v = tet_volumes(x + dx, ien)
            bad = (v * v0 <= 0.0) | (np.abs(v) < 0.3 * np.abs(v0))
            if not bad.any():
                break
            dx[np.unique(ien[bad].reshape(-1))] *= 0.5
        x = x + dx
    ien = _fix_orientation(x, ien).astype(np.int32)

    nid = np.arange(x.shape[0], dtype=np.int32)
    # outlet triangles: the faces of the Kuhn tets lying in the plane k = nz
    sx, sy = 1, nx + 1
    base = ((np.arange(ny)[:, None] * sy + np.arange(nx)[None, :] * sx).reshape(-1) + nz * (nx + 1) * (ny + 1))
    t1 = np.stack([base, base + sx, base + sx + sy], axis=1)
    t2 = np.stack([base, base + sx + sy, base + sy], axis=1)
    tris = np.concatenate([t1, t2]).astype(np.int32)       # counter-clockwise seen from +z => outward normal +z
    faces = {
        "inlet": dict(nodes=nid[on_in], tris=None),
        "outlet": dict(nodes=nid[on_out], tris=tris),
        "wall": dict(nodes=nid[on_wall], tris=None),
    }
    return Mesh(x=np.ascontiguousarray(x), ien=np.ascontiguousarray(ien), faces=faces, shape=(nx, ny, nz))


def block_mesh(n: int, elem: str = "hex", length: float = 1.0, jitter: float = 0.1, seed: int = 4321, curve: float = 0.03) -> Mesh:
    """Cube [0,L]^3 of n^3 HEX8 (reference node order, nn_elem_gnn.h:732: bottom face counter-clockwise, then the
    top face), its 6-tet Kuhn split ("tet") or the quadratic version of that split ("tet10"); interior nodes jittered
    by jitter*h*U(-1,1).  Faces X0..Z1 = node lists."""
    h = length / n
    g = np.arange(n + 1) * h
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")            # node id = i + (n+1) j + (n+1)^2 k
    x = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    rng = np.random.default_rng(seed)
    interior = np.all((x > 1e-12) & (x < length - 1e-12), axis=1)
    x[interior] += jitter * h * rng.uniform(-1.0, 1.0, size=(int(interior.sum()), 3))
    if elem == "hex":
        sx, sy, sz = 1, n + 1, (n + 1) ** 2
        i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        base = (i * sx + j * sy + k * sz).transpose(2, 1, 0).reshape(-1).astype(np.int64)
        off = [0, sx, sx + sy, sy, sz, sx + sz, sx + sy + sz, sy + sz]
        ien = np.stack([base + o for o in off], axis=1).astype(np.int32)
    elif elem == "tet":
        ien = _fix_orientation(x, _kuhn_tets(n, n, n)).astype(np.int32)
    elif elem == "tet10":
        # quadratic tets: the Kuhn split + one node per edge in the reference's order (nn_elem_gnn.h:1256-1268:
        # nodes 4..9 sit on the edges 0-1, 1-2, 0-2, 0-3, 1-3, 2-3); mid-edge nodes of interior edges are moved off
        # the chord by curve*h*U(-1,1), so the elements are genuinely curved (xXi2 != 0 in gn_nxx)
        t4 = _fix_orientation(x, _kuhn_tets(n, n, n)).astype(np.int64)
        pairs = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
        nv = x.shape[0]
        ek = np.stack([np.minimum(t4[:, a], t4[:, b]) * nv + np.maximum(t4[:, a], t4[:, b]) for a, b in pairs], axis=1)
        uniq, inv = np.unique(ek.reshape(-1), return_inverse=True)
        ea, eb = uniq // nv, uniq % nv
        xm = 0.5 * (x[ea] + x[eb])
        inner = interior[ea] | interior[eb]
        xm[inner] += curve * h * rng.uniform(-1.0, 1.0, size=(int(inner.sum()), 3))
        x = np.concatenate([x, xm])
        ien = np.concatenate([t4, nv + inv.reshape(-1, 6)], axis=1).astype(np.int32)
    else:
        raise ValueError(elem)
    nid = np.arange(x.shape[0], dtype=np.int32)
    faces = {}
    for ax, nm in enumerate("XYZ"):
        faces[nm + "0"] = dict(nodes=nid[np.abs(x[:, ax]) < 1e-12], tris=None)
        faces[nm + "1"] = dict(nodes=nid[np.abs(x[:, ax] - length) < 1e-12], tris=None)
    return Mesh(x=np.ascontiguousarray(x), ien=np.ascontiguousarray(ien), faces=faces, shape=(n, n, n))


# local faces of the volume elements, in an order that is a valid face element of the reference (nn_elem_gnn.h:1536-1610:
# QUD4 nodes cyclic; TRI6 = corners 0,1,2 then the mid-edge nodes of 0-1, 1-2, 0-2)
_TET_FACES = [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)]
_TET10_MID = {(0, 1): 4, (1, 2): 5, (0, 2): 6, (0, 3): 7, (1, 3): 8, (2, 3): 9}
_HEX_FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]


def face_elements(mesh: Mesh, on_face: np.ndarray):
    """Boundary-face mesh of the nodes flagged by ``on_face`` (bool per node): every local face of a volume element whose
    nodes are all flagged.  Returns (IENb (nElb, eNoNb) int32, gE (nElb,) int32) = lFa.IEN and lFa.gE of the reference
    (ComMod.h faceType): TRI3 for TET4, QUD4 for HEX8, TRI6 for TET10."""
    ien = mesh.ien
    eNoN = ien.shape[1]
    if eNoN == 8:
        loc = [list(f) for f in _HEX_FACES]
    elif eNoN == 4:
        loc = [list(f) for f in _TET_FACES]
    else:
        mid = lambda a, b: _TET10_MID[(min(a, b), max(a, b))]
        loc = [[i, j, k, mid(i, j), mid(j, k), mid(i, k)] for i, j, k in _TET_FACES]
    out_i, out_e = [], []
    for f in loc:
        nodes = ien[:, f]
        sel = np.nonzero(on_face[nodes].all(axis=1))[0]
        out_i.append(nodes[sel]); out_e.append(sel)
    IENb = np.concatenate(out_i).astype(np.int32)
    gE = np.concatenate(out_e).astype(np.int32)
    order = np.argsort(gE, kind="stable")                  # face elements in the order of their parents
    return np.ascontiguousarray(IENb[order]), np.ascontiguousarray(gE[order])


def block_state(mesh: Mesh, length: float = 1.0, amp: float = 0.05, noise: float = 0.01, seed: int = 2026, tDof: int = 3, s: int = 0):
    """Displacement state of SURVEY.md par. 8d for the solid block: d = amp*L*sin field + noise, Ag/Yg random."""
    x = mesh.x
    n = x.shape[0]
    rng = np.random.default_rng(seed)
    k = np.pi / length
    Dg = np.zeros((n, tDof)); Ag = np.zeros((n, tDof)); Yg = np.zeros((n, tDof))
    Dg[:, s + 0] = amp * length * np.sin(k * x[:, 0]) * np.cos(k * x[:, 1])
    Dg[:, s + 1] = amp * length * np.sin(k * x[:, 1]) * np.cos(k * x[:, 2])
    Dg[:, s + 2] = -amp * length * np.sin(k * x[:, 2]) * np.cos(k * x[:, 0])
    Dg[:, s:s + 3] += noise * amp * length * rng.standard_normal((n, 3))
    Yg[:, s:s + 3] = 0.1 * rng.standard_normal((n, 3))
    Ag[:, s:s + 3] = 10.0 * rng.standard_normal((n, 3))
    Bf = 0.5 * rng.standard_normal((n, 3))
    return Ag, Yg, Dg, Bf


def gen_alpha2(ro_inf: float = 0.5):
    """Generalised-alpha constants for a second-order (solid) equation (S/initialize.cpp:424-463)."""
    am = (2.0 - ro_inf) / (1.0 + ro_inf)
    af = 1.0 / (1.0 + ro_inf)
    beta = 0.25 * (1.0 + am - af) ** 2
    gam = 0.5 + am - af
    return am, af, gam, beta


def csr_pattern(ien: np.ndarray, nNo: int):
    """Node-graph CSR with sorted columns and the diagonal present.

    Same pattern lhsa_ns::lhsa builds from IEN when no undeformed-Neumann face rewires idMap
    (Code/Source/solver/lhsa.cpp:153-380): rowPtr(tnNo+1), colPtr(nnz), 0-based, 32-bit.
    """
    eNoN = ien.shape[1]
    rows = np.repeat(ien.astype(np.int64), eNoN, axis=1).reshape(-1)
    cols = np.tile(ien.astype(np.int64), (1, eNoN)).reshape(-1)
    key = np.unique(rows * nNo + cols)
    r = (key // nNo).astype(np.int32)
    c = (key % nNo).astype(np.int32)
    rowPtr = np.zeros(nNo + 1, dtype=np.int32)
    np.add.at(rowPtr, r + 1, 1)
    rowPtr = np.cumsum(rowPtr, dtype=np.int64).astype(np.int32)
    return rowPtr, c


def face_normal_integral(x: np.ndarray, tris: np.ndarray, nodes: np.ndarray) -> np.ndarray:
    """val(:,a) = int N_a n dGamma for a face of linear triangles (S/baf_ini.cpp:746-770).

    For TRI3 the 3-point rule integrates N_a exactly: area * n / 3 per vertex.  Returned (nNodes, 3)
    in the order of ``nodes``.
    """
    p0, p1, p2 = x[tris[:, 0]], x[tris[:, 1]], x[tris[:, 2]]
    an = 0.5 * np.cross(p1 - p0, p2 - p0)                  # area-weighted normal
    acc = np.zeros_like(x)
    for a in range(3):
        np.add.at(acc, tris[:, a], an / 3.0)
    return np.ascontiguousarray(acc[nodes])


def pipe_state(mesh: Mesh, radius: float = 1.0, length: float = 10.0, umax: float = 20.0,
               dp: float = 100.0, noise: float = 0.01, seed_y: int = 2024, seed_a: int = 2025,
               tDof: int = 4):
    """Assembly-parity state of SURVEY.md §8d: Poiseuille profile + 1% Gaussian noise in Yg, linear
    pressure drop + noise, Ag ~ 10*N(0,1), Bf = 0.  Returns (Ag, Yg, Bf) as (nNo, tDof)/(nNo, 3)."""
    x = mesh.x
    n = x.shape[0]
    rng_y = np.random.default_rng(seed_y)
    rng_a = np.random.default_rng(seed_a)
    r2 = (x[:, 0] ** 2 + x[:, 1] ** 2) / radius ** 2
    Yg = np.zeros((n, tDof))
    Yg[:, 2] = umax * np.clip(1.0 - r2, 0.0, None)
    Yg[:, :3] += noise * umax * rng_y.standard_normal((n, 3))
    Yg[:, 3] = dp * (1.0 - x[:, 2] / length) + noise * dp * rng_y.standard_normal(n)
    Ag = np.zeros((n, tDof))
    Ag[:, :4] = 10.0 * rng_a.standard_normal((n, 4))
    Bf = np.zeros((n, 3))
    return Ag, Yg, Bf


def gen_alpha(ro_inf: float = 0.5):
    """Generalised-alpha constants for a first-order (fluid) equation (S/initialize.cpp:424-463)."""
    am = 0.5 * (3.0 - ro_inf) / (1.0 + ro_inf)
    af = 1.0 / (1.0 + ro_inf)
    gam = 0.5 + am - af
    return am, af, gam
