"""GPU: the CSR value reductions agree bit for bit.

Three reductions produce the CSR values from the per-element stiffness scratch (the reference's slot-map scatter,
SparseAssemblyNative.h:32-45 / _MassIntegrand_.h:115-166): the shared-memory row-buffer kernels (fl_pattern.cu), the register-resident
slot-owner gather (fl_gather.cuh, fl_set_option 3 = 1) and, for LinearElastic on tet10 / hex8, the curve-ordered assembly of
fl_stream.cu (fl_set_option 4 = 1 two-pass, 2 concurrent on two streams, 3 the concurrent kernels one after the other).  All of them
add the visits of a node in ascending ORIGINAL element number starting from +0.0, so the results must be identical, not merely close.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_gpu_parity import ELEC, _cases, _load, _material  # noqa: E402


def _select(h, opt):
    """0: element order, row-buffer kernels; 1: element order, register gather; 2: curve order + register gather for every shape the
    gather is instantiated for (FL_CURVE_ALL; by default only the shapes where it was measured faster); None: defaults."""
    os.environ.pop("FL_CURVE_ALL", None)
    if opt is None:
        h.set_option(3, 2); h.set_option(4, 1)
        return
    if opt == 2:
        os.environ["FL_CURVE_ALL"] = "1"
    h.set_option(3, min(opt, 1)); h.set_option(4, 1 if opt == 2 else 0)


def _supported_by_register_gather(nvar, npe):
    return (nvar == 2 and npe in (3, 4, 6, 9)) or (nvar == 3 and npe in (4, 8, 10)) or (nvar == 4 and npe in (4, 8, 10, 27))


@pytest.mark.parametrize("key", _cases())
def test_register_gather_equals_row_buffer_gather_on_the_golden_cases(key):
    """Every reference-generated case: same CSR values from both reductions (shapes the register gather does not cover must fall
    back silently to the row-buffer kernels and still agree)."""
    from florence_b200 import backend
    from oracle import oracle as orc
    c = _load(key)
    matname = key.split("_", 3)[3]
    num = orc.MATERIAL_NUMBERS[matname]
    ndim = c["points"].shape[1]
    electro = matname in ELEC
    nvar, form, update = ndim + (1 if electro else 0), (1 if electro else 0), int(c["update"])
    h = backend.AssemblyHandle(c["points"], c["elements"], c["Jm"], c["AllGauss"], c["Bases"])
    mat = _material(backend, num, c["prm"])
    h.build_pattern(nvar)
    out = {}
    for opt in (0, 1, 2):      # row buffer, register gather (element order), K_e along the curve + register gather (any material)
        _select(h, opt)
        V, T = h.assemble_implicit(c["Eulerx"], c["Eulerp"], mat, form, update, mode="csr")
        out[opt] = (V.clone(), T.clone())
    _select(h, None)
    for opt in (1, 2):
        assert torch.equal(out[0][0], out[opt][0]) and torch.equal(out[0][1], out[opt][1]), opt
    h.close()


@pytest.mark.parametrize("kind,p,n,nvar", [("tri", 1, 9, 2), ("tri", 2, 7, 2), ("quad", 1, 8, 2), ("quad", 2, 6, 2), ("tet", 1, 5, 3),
                                           ("hex", 1, 6, 3), ("tet", 2, 4, 3), ("tet", 1, 4, 4), ("hex", 1, 4, 4), ("tet", 2, 3, 4),
                                           ("hex", 2, 3, 4)])
def test_register_gather_on_larger_meshes(kind, p, n, nvar):
    """Meshes with interior nodes of full valence (tet10 vertices: 65 neighbours = three groups of 32 slots, 24 elements = several
    steps) for every instantiated (nvar, nodes per element): NeoHookean / electro-mechanics 108, both reductions, identical values."""
    from florence_b200 import backend, mesh as flmesh
    pts, els = flmesh.make_mesh(kind, n, p)
    B, Jm, AG = flmesh.tables(kind, p)
    ndim = pts.shape[1]
    assert _supported_by_register_gather(nvar, els.shape[1])
    x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.05, seed=3)
    xp = None
    if nvar == ndim:
        mat, form = backend.make_material(1, 1.0, mu=3.0, lamb=7.0), 0
    else:
        mat, form = backend.make_material(8, 1.0, mu1=2.0, mu2=1.5, lamb=6.0, eps_1=3.0, eps_2=2.0), 1
        xp = 0.1 * torch.sin(5.0 * pts.sum(1))
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    h.build_pattern(nvar)
    out = {}
    for opt in (0, 1, 2, 1, 2):
        _select(h, opt)
        V, T = h.assemble_implicit(x, xp, mat, form, True, mode="csr")
        if opt in out:
            assert torch.equal(V, out[opt][0]) and torch.equal(T, out[opt][1]), "bit-reproducible"
        out[opt] = (V.clone(), T.clone())
    _select(h, None)
    assert out[0][0].abs().max() > 0
    for opt in (1, 2):
        assert torch.equal(out[0][0], out[opt][0]) and torch.equal(out[0][1], out[opt][1]), opt
    h.close()


@pytest.mark.parametrize("kind,p,n,nel", [("tet", 2, 2, 1), ("tet", 2, 2, 7), ("tet", 2, 3, None), ("tet", 2, 6, None), ("tet", 2, 11, None),
                                          ("hex", 1, 3, 9), ("hex", 1, 7, None), ("hex", 1, 12, None)])
def test_curve_ordered_assembly_equals_element_order_assembly(kind, p, n, nel):
    """LinearElastic tet10 / hex8, CSR mode: K_e stored along the Morton curve (two-pass, concurrent on two streams, and the
    concurrent kernels run one after the other) gives the same V and T, bit for bit, as the element-order two-pass path -- ragged
    groups, a single element, meshes larger than one wave of element groups, both detJ rules, and repeated calls (the progress
    flags carry the epoch of the call: a second call must not see the first call's flags)."""
    from florence_b200 import backend, mesh as flmesh
    os.environ["FL_STREAM_CHECK"] = "1"       # surface a reduction warp that gave up waiting as an error
    pts, els = flmesh.make_mesh(kind, n, p)
    if nel is not None:
        els = els[:nel]
    B, Jm, AG = flmesh.tables(kind, p)
    x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.05, seed=5)
    mat = backend.make_material(10, 1.0, mu=3.0, lamb=7.0)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    h.build_pattern(3)
    h.set_option(2, 2)                         # warp-autonomous element kernel for hex8 as well
    for update in (True, False):
        h.set_option(4, 0)
        Vr, Tr = h.assemble_implicit(x, None, mat, 0, update, mode="csr")
        Vr, Tr = Vr.clone(), Tr.clone()
        for mode in (1, 2, 3, 2):
            h.set_option(4, mode)
            for _ in range(2):
                V, T = h.assemble_implicit(x, None, mat, 0, update, mode="csr")
                assert torch.equal(V, Vr) and torch.equal(T, Tr), "mode %d" % mode
    h.set_option(4, 1)
    h.set_option(2, 1)
    h.close()


def test_register_gather_on_a_fan_of_tetrahedra():
    """One node shared by 150 tet4 elements with ~190 neighbours: several items per node (more than 96 slots), records fetched in
    more than one batch of 32 visits, many steps per item -- against the row-buffer reduction (bit for bit) and the oracle."""
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    rng = np.random.default_rng(7)
    nnode, nelem = 200, 150
    pts = rng.random((nnode, 3))
    els = np.zeros((nelem, 4), dtype=np.int64)
    for e in range(nelem):
        els[e, 0] = 0
        els[e, 1:] = rng.choice(np.arange(1, nnode), 3, replace=False)
    for e in range(nelem):                      # positive orientation, so that the kinematics are the ordinary ones
        X = pts[els[e]]
        if np.linalg.det(X[1:] - X[0]) < 0:
            els[e, [2, 3]] = els[e, [3, 2]]
    B, Jm, AG = flmesh.tables("tet", 1)
    x = pts + 0.01 * np.sin(5 * pts + 0.3)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    mat = backend.make_material(10, 1.0, mu=3.0, lamb=7.0)
    prm = orc.params(mu=3.0, lamb=7.0)
    pat = orc.sparsity_pattern(els, nnode, 3)
    Vo, To = orc.assemble_implicit(pts, els, x, None, Jm, AG, 3, 6, 1, prm, 10, mode="csr", pattern=pat)
    h.build_pattern(3)
    out = {}
    for opt in (0, 1):
        h.set_option(3, opt)
        V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
        out[opt] = V.clone()
        assert np.abs(V.cpu().numpy() - Vo).max() <= 1e-10 * np.abs(Vo).max()
        assert np.linalg.norm(T.cpu().numpy() - To) <= 1e-11 * np.linalg.norm(To)
    assert torch.equal(out[0], out[1])
    h.close()


def test_curve_ordered_assembly_with_shuffled_renumbered_and_repeated_elements():
    """Nothing may depend on the mesh's own ordering: tet10 mesh with randomly renumbered nodes, shuffled elements and every element
    listed five times (vertex nodes are then visited 120 times: several batches of records, twenty steps) -- all curve-ordered modes
    against the element-order path (bit for bit) and the oracle."""
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    os.environ["FL_STREAM_CHECK"] = "1"
    rng = np.random.default_rng(11)
    pts, els = flmesh.box_tet_mesh(3, 3, 3, p=2)
    pts, els = pts.numpy(), els.numpy()
    perm = rng.permutation(pts.shape[0])                  # new id of old node k
    P = np.empty_like(pts); P[perm] = pts
    E = perm[els]
    E = np.concatenate([E] * 5)[rng.permutation(5 * els.shape[0])]
    B, Jm, AG = flmesh.tables("tet", 2)
    x = P + 0.004 * np.sin(7 * P + 1.0)
    h = backend.AssemblyHandle(P, E, Jm, AG, B)
    mat = backend.make_material(10, 1.0, mu=3.0, lamb=7.0)
    prm = orc.params(mu=3.0, lamb=7.0)
    pat = orc.sparsity_pattern(E, P.shape[0], 3)
    Vo, To = orc.assemble_implicit(P, E, x, None, Jm, AG, 3, 6, 1, prm, 10, mode="csr", pattern=pat)
    h.build_pattern(3)
    h.set_option(4, 0)
    Vr, Tr = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
    Vr, Tr = Vr.clone(), Tr.clone()
    assert np.abs(Vr.cpu().numpy() - Vo).max() <= 1e-10 * np.abs(Vo).max()
    assert np.linalg.norm(Tr.cpu().numpy() - To) <= 1e-11 * np.linalg.norm(To)
    for mode in (1, 2, 3):
        h.set_option(4, mode)
        V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
        assert torch.equal(V, Vr) and torch.equal(T, Tr), "mode %d" % mode
    h.set_option(4, 1)
    h.close()
