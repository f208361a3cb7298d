"""Block-CSR pattern construction on the device (SURVEY.md par. 8(f) row 3): b200_pattern_* against lhsa_ns::lhsa
(Code/Source/solver/lhsa.cpp:153) -- integer work, exact equality of rowPtr / colPtr."""
import time

import numpy as np
import pytest

from svfsiplus_b200 import backend as B
from svfsiplus_b200 import mesh as M

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("elem,n", [("tet", 5), ("hex", 5), ("tet10", 3)])
def test_pattern_equals_reference_lhsa(elem, n):
    m = M.block_mesh(n, elem)
    be = B.Backend(0)
    rp, cp = be.pattern(m.nNo, [m.ien])
    rp2, cp2 = M.csr_pattern(m.ien, m.nNo)                 # equals the reference's lhsa (tests/test_cpu_oracle.py)
    assert np.array_equal(rp, rp2) and np.array_equal(cp, cp2)
    from oracle import ref
    if ref.available():
        ra = ref.RefAssembly(m.x, m.ien)
        rp3, cp3 = ra.csr()
        ra.close()
        assert np.array_equal(rp, rp3) and np.array_equal(cp, cp3)
    be.close()


def test_pattern_of_two_meshes_and_ragged_ids():
    """Two meshes sharing an interface (the reference loops over com_mod.msh) and isolated nodes (empty rows)."""
    m = M.pipe_mesh(5, 4, 7)
    half = m.nEl // 2
    be = B.Backend(0)
    rp, cp = be.pattern(m.nNo + 3, [m.ien[:half], m.ien[half:]])       # three nodes no element touches
    rp2, cp2 = M.csr_pattern(m.ien, m.nNo + 3)
    assert np.array_equal(rp, rp2) and np.array_equal(cp, cp2)
    assert rp[-1] == rp[-4]                                             # the isolated nodes own empty rows
    be.close()


def test_pattern_large_mesh_properties():
    """1.33 M tets (the size of test_large_mesh_properties): sortedness, diagonal present, symmetry of the pattern,
    nnz = nNo + 2 * edges, and the device builder is much faster than the serial host construction."""
    m = M.pipe_mesh(48, 48, 96)
    be = B.Backend(0)
    t0 = time.time()
    rp, cp = be.pattern(m.nNo, [m.ien])
    t_dev = time.time() - t0
    assert rp[0] == 0 and rp[-1] == len(cp)
    rows = np.repeat(np.arange(m.nNo), np.diff(rp))
    assert (np.diff(cp)[np.diff(rows) == 0] > 0).all()                  # strictly increasing inside every row
    diag = np.zeros(m.nNo, bool); diag[rows[cp == rows]] = True
    assert diag.all()
    key = rows.astype(np.int64) * m.nNo + cp
    keyT = cp.astype(np.int64) * m.nNo + rows
    assert np.array_equal(np.sort(key), np.sort(keyT))                  # structurally symmetric
    e = np.sort(m.ien[:, [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]].reshape(-1, 2), axis=1)
    nedges = len(np.unique(e[:, 0].astype(np.int64) * m.nNo + e[:, 1]))
    assert len(cp) == m.nNo + 2 * nedges
    print(f"device pattern of {m.nEl} tets: {t_dev:.3f} s (incl. the upload of IEN and the download of the CSR)")
    be.close()
