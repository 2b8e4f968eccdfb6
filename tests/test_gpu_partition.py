"""GPU: the partitioned paths of SURVEY.md 8(e) on one device, rank by rank -- the CSR row blocks emitted by the ranks of an
implicit run assemble the single-domain matrix bit for bit, and the per-rank internal forces summed over the interface equal the
single-domain force.  (The process-group side -- NCCL / gloo exchange -- is covered by tests/test_partition_gloo.py.)"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,p,n,matnum,world", [("tet", 2, 4, 10, 2), ("hex", 1, 5, 1, 3), ("hex", 2, 3, 2, 2)])
def test_implicit_row_blocks_reassemble_the_global_matrix(kind, p, n, matnum, world):
    from florence_b200 import backend, mesh as flmesh, partition
    pts, els = (flmesh.box_tet_mesh if kind == "tet" else flmesh.box_hex_mesh)(n, n, n, p=p)
    B, Jm, AG = flmesh.tables(kind, p)
    x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.02, seed=4)
    prm = dict(mu=1e5, lamb=1.5e5) if matnum != 2 else dict(mu1=6e4, mu2=4e4, lamb=1.5e5)
    mat = backend.make_material(matnum, 1.0, **prm)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    h.build_pattern(3)
    V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
    indices, indptr = h.sparsity_pattern(3)
    V, indices, indptr = V.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()
    h.close()
    N = 3 * pts.shape[0]
    seen = np.zeros(N, bool)
    for rank in range(world):
        part = partition.row_partition(pts.numpy(), els.numpy(), rank, world)
        hl = backend.AssemblyHandle(part.points, part.elements, Jm, AG, B)
        hl.build_pattern(3)
        xl = x[part.node_map]
        Vl, Tl = hl.assemble_implicit(xl, None, mat, 0, True, mode="csr")
        il, pl = hl.sparsity_pattern(3)
        rows, ptr_b, cols, vals = part.owned_rows(Vl.cpu().numpy(), il.cpu().numpy(), pl.cpu().numpy(), 3)
        hl.close()
        assert not seen[rows].any()
        seen[rows] = True
        for k, r in enumerate(rows):
            s, e = indptr[r], indptr[r + 1]
            assert np.array_equal(cols[ptr_b[k]:ptr_b[k + 1]], indices[s:e])
            assert np.array_equal(vals[ptr_b[k]:ptr_b[k + 1]], V[s:e]), "row %d of rank %d differs" % (r, rank)
    assert seen.all()


@pytest.mark.parametrize("p,n,world", [(1, 6, 2), (2, 3, 3)])
def test_explicit_partial_forces_sum_to_the_global_force(p, n, world):
    from florence_b200 import backend, mesh as flmesh, partition
    pts, els = flmesh.box_hex_mesh(n, n, n, p=p)
    B, Jm, AG = flmesh.tables("hex", p)
    x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.02, seed=8)
    mat = backend.make_material(1, 1100.0, mu=4e5, lamb=2e6)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    T = h.assemble_explicit(x, None, mat, 0).cpu().numpy().reshape(-1, 3)
    M = h.assemble_mass(1100.0, 3, "lumped").cpu().numpy().reshape(-1, 3)
    h.close()
    Tsum, Msum = np.zeros_like(T), np.zeros_like(M)
    for rank in range(world):
        part = partition.partition_mesh(pts.numpy(), els.numpy(), rank, world)
        hl = backend.AssemblyHandle(part.points, part.elements, Jm, AG, B)
        gl = part.node_map.numpy()
        Tsum[gl] += hl.assemble_explicit(x[part.node_map], None, mat, 0).cpu().numpy().reshape(-1, 3)
        Msum[gl] += hl.assemble_mass(1100.0, 3, "lumped").cpu().numpy().reshape(-1, 3)
        # the pack / unpack kernels of the interface exchange: packing then adding back doubles the shared entries only
        import ctypes as C
        from florence_b200._lib import check
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for other, ids in part.neighbours.items():
            ids_d = ids.to(hl.device)
            v = torch.arange(part.points.shape[0] * 3, dtype=torch.float64, device=hl.device)
            buf = torch.empty(ids.numel() * 3, dtype=torch.float64, device=hl.device)
            check(hl.lib.fl_pack_nodes(C.c_void_p(v.data_ptr()), C.c_void_p(ids_d.data_ptr()), ids_d.numel(), 3, C.c_void_p(buf.data_ptr()), st))
            assert torch.equal(buf.view(-1, 3), v.view(-1, 3)[ids_d.long()])
            w = v.clone()
            check(hl.lib.fl_unpack_add_nodes(C.c_void_p(w.data_ptr()), C.c_void_p(ids_d.data_ptr()), ids_d.numel(), 3, C.c_void_p(buf.data_ptr()), st))
            expect = v.clone().view(-1, 3)
            expect[ids_d.long()] *= 2
            assert torch.equal(w.view(-1, 3), expect)
            check(hl.lib.fl_scatter_nodes(C.c_void_p(w.data_ptr()), C.c_void_p(ids_d.data_ptr()), ids_d.numel(), 3, C.c_void_p(buf.data_ptr()), st))
            assert torch.equal(w, v)
        hl.close()
    assert np.abs(Tsum - T).max() <= 1e-12 * np.abs(T).max()
    assert np.abs(Msum - M).max() <= 1e-13 * np.abs(M).max()


@pytest.mark.parametrize("kind,p,n,world", [("hex", 2, 4, 3), ("tet", 2, 3, 4)])
def test_device_localisation_equals_host_localisation(kind, p, n, world):
    from florence_b200 import mesh as flmesh, partition
    pts, els = (flmesh.box_tet_mesh if kind == "tet" else flmesh.box_hex_mesh)(n, n, n, p=p)
    dev = torch.device("cuda:0")
    for rank in range(world):
        ph = partition.partition_mesh(pts.numpy(), els.numpy(), rank, world)
        pd = partition.partition_mesh(pts.to(dev), els.to(dev), rank, world)
        assert pd.points.is_cuda and pd.elements.is_cuda
        assert torch.equal(pd.points.cpu(), ph.points) and torch.equal(pd.elements.cpu(), ph.elements.long())
        assert torch.equal(pd.node_map.cpu(), ph.node_map.long())
        assert sorted(pd.neighbours) == sorted(ph.neighbours)
        for r in ph.neighbours:
            assert torch.equal(pd.neighbours[r].cpu(), ph.neighbours[r])
