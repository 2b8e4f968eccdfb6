"""The reference's OWN native code for the integer side of the path (oracle/_ref/libflorence_ref.so, compiled from
/root/reference by `make -C oracle ref`) against the oracle restatement (CPU) and the device kernels (GPU), bit for bit."""
import numpy as np
import pytest

from florence_b200 import mesh as flmesh
from oracle import oracle as orc

needs_ref = pytest.mark.skipif(not (orc.ref_available() or orc.build_ref()), reason="oracle/_ref not built (no /root/reference here)")

MESHES = [("tet", 2, (5, 4, 6), 3), ("hex", 2, (3, 4, 3), 3), ("hex", 1, (6, 5, 4), 4), ("quad", 2, (7, 5), 2), ("tri", 2, (6, 6), 3), ("hex", 3, (2, 2, 2), 1)]


@needs_ref
@pytest.mark.parametrize("etype,p,n,nvar", MESHES)
def test_oracle_pattern_equals_reference_native(etype, p, n, nvar):
    pts, els = flmesh.make_mesh(etype, n, p)
    rng = np.random.default_rng(1)
    els = els.numpy()[rng.permutation(els.shape[0])]          # element order must not matter for the pattern
    ref = orc.ref_sparsity_pattern(els, pts.shape[0], nvar)
    mine = orc.sparsity_pattern(els, pts.shape[0], nvar)
    for a, b in zip(ref, mine):
        assert a.dtype == b.dtype == np.int32 and np.array_equal(a, b)
    # the reference's CSR scatter through its slot maps equals a scipy COO sum of the same element matrices
    from scipy.sparse import coo_matrix
    ndof = nvar * els.shape[1]
    Ke = rng.standard_normal((els.shape[0], ndof * ndof))
    V = orc.ref_csr_scatter(Ke, ref[2], ref[3], ref[0].shape[0])
    dof = (els[:, :, None] * nvar + np.arange(nvar)[None, None, :]).reshape(els.shape[0], ndof)
    I = np.repeat(dof, ndof, axis=1).ravel()
    J = np.tile(dof, (1, ndof)).ravel()
    K = coo_matrix((Ke.ravel(), (I, J)), shape=(nvar * pts.shape[0],) * 2).tocsr()
    K.sort_indices()
    assert np.array_equal(K.indices, ref[0]) and np.array_equal(K.indptr, ref[1])
    assert np.abs(K.data - V).max() <= 1e-13 * np.abs(V).max()


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("etype,p,n,nvar", [("tet", 2, (12, 12, 12), 3), ("hex", 2, (8, 8, 8), 4), ("quad", 2, (40, 30), 2)])
def test_device_pattern_and_slot_maps_equal_reference_native(etype, p, n, nvar):
    from florence_b200 import backend
    pts, els = flmesh.make_mesh(etype, n, p)
    B, Jm, AG = flmesh.tables(etype, p)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    out = h.sparsity_pattern(nvar, with_data_indices=True)
    ref = orc.ref_sparsity_pattern(els.numpy(), pts.shape[0], nvar)
    for a, b in zip(ref, out):
        assert np.array_equal(a, b.cpu().numpy())
    h.close()
