"""Host-side checks (CPU) of the synthetic mesh / table generator against the reference's FunctionSpace fixtures and against
identities every function space satisfies (Florence/Utils/debug.py:16-54): sum_a N_a = 1, sum_a grad N_a = 0, sum w = |parent|."""
import numpy as np
import pytest

from florence_b200 import mesh as flmesh
from oracle import oracle as orc

CASES = [("hex", 1), ("hex", 2), ("hex", 3), ("hex", 4), ("quad", 1), ("quad", 2), ("tet", 1), ("tet", 2), ("tri", 1), ("tri", 2)]


@pytest.mark.parametrize("etype,p", CASES)
def test_partition_of_unity_and_weights(etype, p):
    Bases, Jm, AG = flmesh.tables(etype, p)
    assert np.allclose(Bases.sum(0), 1.0, atol=1e-13)
    assert np.allclose(Jm.sum(1), 0.0, atol=1e-12)
    vol = {"hex": 8.0, "quad": 4.0, "tet": 1.0 / 6.0, "tri": 0.5}[etype]
    assert abs(AG.sum() - vol) < 1e-13


@pytest.mark.parametrize("etype,p", CASES)
def test_element_laplacian_spectrum_matches_reference_tables(golden, etype, p):
    """K_e of the Laplacian on one straight-sided element is invariant to node and Gauss-point ordering up to a permutation:
    its eigenvalues computed with our tables equal those computed with the reference's FunctionSpace tables."""
    g = golden.tables
    Jm_ref, AG_ref = g["tab_%s%d_Jm" % (etype, p)], g["tab_%s%d_AllGauss" % (etype, p)]
    pts_ref, els_ref = g["tab_%s%d_points" % (etype, p)], g["tab_%s%d_elements" % (etype, p)][:1]
    ndim = pts_ref.shape[1]
    Iref, Jref, Vref = orc.assemble_laplacian(pts_ref, els_ref, Jm_ref, AG_ref, np.eye(ndim), True)
    npe = els_ref.shape[1]
    Kref = Vref.reshape(npe, npe)
    Bases, Jm, AG = flmesh.tables(etype, p)
    if etype in ("hex", "quad"):
        # the same physical element with our generator: a 1-cell mesh with the reference's box dimensions
        L = tuple(pts_ref.max(0) - pts_ref.min(0))
        pts, els = flmesh.make_mesh(etype, (1,) * ndim, p, lengths=L)
        pts, els = pts.numpy(), els.numpy()
    else:
        # simplices: the reference's hex->tet split differs from our Kuhn split, so rebuild ITS first element in our node
        # ordering (vertices = its first ndim+1 local nodes, then our edge midpoints)
        V = pts_ref[els_ref[0, :ndim + 1]]
        if np.linalg.det(V[1:] - V[0]) < 0:
            V[[0, 1]] = V[[1, 0]]
        edges = flmesh.TET10_EDGES if ndim == 3 else flmesh.TRI6_EDGES
        pts = V if p == 1 else np.concatenate([V, np.array([(V[i] + V[j]) / 2 for i, j in edges])])
        els = np.arange(pts.shape[0])[None, :]
    vols = []
    for e in range(els.shape[0]):
        I, J, V = orc.assemble_laplacian(pts, els[e:e + 1], Jm, AG, np.eye(ndim), True)
        vols.append(V.reshape(npe, npe))
    ev_ref = np.sort(np.linalg.eigvalsh(Kref))
    # simplices: the reference's first element is congruent to one of our Kuhn simplices
    ok = any(np.allclose(np.sort(np.linalg.eigvalsh(K)), ev_ref, rtol=1e-9, atol=1e-10 * abs(ev_ref).max()) for K in vols)
    assert ok


def test_mesh_sizes_of_the_benchmark_configs():
    pts, els = flmesh.box_hex_mesh(4, 3, 2, p=2)
    assert pts.shape == (9 * 7 * 5, 3) and els.shape == (24, 27)
    pts, els = flmesh.box_tet_mesh(3, 3, 3, p=2)
    assert pts.shape == (7 ** 3, 3) and els.shape == (6 * 27, 10)
    assert len(np.unique(els.numpy())) == 7 ** 3  # every fine-grid point is a tet10 node ((2N+1)^3 nodes)


@pytest.mark.parametrize("etype,p", CASES)
def test_all_elements_positively_oriented_and_volume(etype, p):
    n = 2
    pts, els = flmesh.make_mesh(etype, n, p)
    Bases, Jm, AG = flmesh.tables(etype, p)
    X = pts.numpy()[els.numpy()]                       # (nelem, npe, d)
    J = np.einsum("kag,eal->egkl", Jm, X)
    det = np.linalg.det(J)
    assert (det > 0).all()
    assert abs((det * AG.ravel()[None, :]).sum() - 1.0) < 1e-12


def test_rigid_translation_gives_zero_internal_force():
    pts, els = flmesh.box_hex_mesh(2, 2, 2, p=2)
    Bases, Jm, AG = flmesh.tables("hex", 2)
    x = pts.numpy() + np.array([0.3, -0.2, 0.1])
    T = orc.assemble_explicit(pts.numpy(), els.numpy(), x, None, Jm, AG, 3, orc.params(mu=4e5, lamb=2e6), 1)
    assert np.abs(T).max() < 1e-6
