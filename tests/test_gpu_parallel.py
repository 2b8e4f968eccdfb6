"""GPU: the partitioned entry points on one device -- launcher drop-ins in serial execution, the device row-block emission, the
space-filling-curve order, and the rank-ordered interface sums with the messages copied by hand between per-rank exchange objects.
(Process groups: tests/test_partition_gloo.py, tests/test_dispatch_host.py on CPU; tests/multigpu_check.py under torchrun.)"""
import os
import subprocess
import sys

import numpy as np
import pytest
from scipy.sparse import csr_matrix

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, os.path.dirname(__file__))


def test_sfc_order_device_equals_host_twin():
    from florence_b200 import mesh as flmesh, partition
    for kind, n, p in (("hex", 9, 1), ("tet", 5, 2), ("hex", 4, 2)):
        pts, els = (flmesh.box_tet_mesh if kind == "tet" else flmesh.box_hex_mesh)(n, n + 1, n - 1, p=p, lengths=(1.0, 0.7, 1.3))
        g = torch.Generator(); g.manual_seed(3)
        els = els[torch.randperm(els.shape[0], generator=g)]
        host = partition.sfc_order(pts.numpy(), els.numpy())
        dev = partition.sfc_order(pts.cuda(), els.cuda()).cpu().numpy()
        assert np.array_equal(host, dev), kind
    # 2-D
    pts, els = flmesh.rect_quad_mesh(7, 5, p=2)
    assert np.array_equal(partition.sfc_order(pts.numpy(), els.numpy()), partition.sfc_order(pts.cuda(), els.cuda()).cpu().numpy())


@pytest.mark.parametrize("kind,p,n,world,order", [("tet", 2, 4, 3, "sfc"), ("hex", 2, 3, 2, None)])
def test_device_row_block_equals_host_slicing(kind, p, n, world, order):
    from florence_b200 import backend, mesh as flmesh, partition
    pts, els = (flmesh.box_tet_mesh if kind == "tet" else flmesh.box_hex_mesh)(n, n, n, p=p)
    B, Jm, AG = flmesh.tables(kind, p)
    x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.02, seed=4)
    mat = backend.make_material(10, 1.0, mu=1e5, lamb=1.5e5)
    nnz_total, rows_total = 0, 0
    for rank in range(world):
        part = partition.row_partition(pts.numpy(), els.numpy(), rank, world, order=order)
        hl = backend.AssemblyHandle(part.points, part.elements, Jm, AG, B)
        hl.build_pattern(3)
        Vl, Tl = hl.assemble_implicit(x[part.node_map], None, mat, 0, True, mode="csr")
        il, pl = hl.sparsity_pattern(3)
        rows, ptr_b, cols, vals = part.owned_rows(Vl.cpu().numpy(), il.cpu().numpy(), pl.cpu().numpy(), 3)
        ip_d, cols_d, vals_d = hl.row_block(3, Vl, part.owned_local, part.node_map)
        assert ip_d.dtype == torch.int64 and cols_d.dtype == torch.int64
        assert np.array_equal(ip_d.cpu().numpy(), ptr_b) and np.array_equal(cols_d.cpu().numpy(), cols)
        assert np.array_equal(vals_d.cpu().numpy(), vals)
        assert np.array_equal(part.global_rows(3).numpy(), rows)
        nnz_total += vals.shape[0]; rows_total += rows.shape[0]
        hl.close()
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    assert nnz_total == h.build_pattern(3) and rows_total == 3 * pts.shape[0]
    h.close()


@pytest.mark.parametrize("world,order", [(3, None), (4, "sfc")])
def test_rank_ordered_interface_sums_are_bit_identical_on_every_sharer(world, order):
    """All ranks of a partitioned explicit force evaluation in one process: per-rank handles and InterfaceExchange objects, the
    NCCL messages replaced by device copies (send buffer of a -> receive buffer of b)."""
    from florence_b200 import backend, mesh as flmesh, partition
    pts, els = flmesh.box_tet_mesh(3, 2, 4, p=2)
    B, Jm, AG = flmesh.tables("tet", 2)
    x = flmesh.perturbed_state(pts, 0.25, 0.02, seed=8)
    mat = backend.make_material(1, 1100.0, mu=4e5, lamb=2e6)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B)
    T = h.assemble_explicit(x, None, mat, 0).cpu().numpy().reshape(-1, 3)
    h.close()
    parts = [partition.partition_mesh(pts.numpy(), els.numpy(), r, world, order=order).interface_first() for r in range(world)]
    hs = [backend.AssemblyHandle(p.points, p.elements, Jm, AG, B) for p in parts]
    exs = [partition.InterfaceExchange(p, 3, hl.device, handle=hl) for p, hl in zip(parts, hs)]
    for p, hl, ex in zip(parts, hs, exs):
        xl = x[p.node_map].to(hl.device)
        nb = p.n_interface_elements
        assert 0 < nb <= hl.nelem
        hl.explicit_forces(xl, mat, 0, nb)                 # interface elements first ...
        hl.gather_pack_nodes(3, ex.U, ex.own)              # ... complete the interface partial sums
        hl.explicit_forces(xl, mat, nb, hl.nelem)
        for r in ex.ranks:
            ex._pack(ex.own, ex.pos[r], ex.send[r])
    for a, ea in enumerate(exs):
        for b in ea.ranks:
            exs[b].recv[a].copy_(ea.send[b])
    copies = {}
    for p, hl, ex in zip(parts, hs, exs):
        ex._sum_ordered()
        Tl = torch.empty(hl.nnode * 3, dtype=torch.float64, device=hl.device)
        hl.gather_nodes(3, Tl)
        ex._scatter(Tl)
        gl = p.node_map.numpy()
        Tn = Tl.cpu().numpy().reshape(-1, 3)
        assert np.abs(Tn - T[gl]).max() <= 1e-12 * np.abs(T).max()
        for i in ex.U.cpu().numpy():
            copies.setdefault(int(gl[i]), []).append(Tn[i].tobytes())
        assert int((ex.iface_slot >= 0).sum()) == ex.n_interface
    assert max(len(v) for v in copies.values()) >= 3
    for node, lst in copies.items():
        assert all(b == lst[0] for b in lst), node
    for hl in hs:
        hl.close()


def test_launchers_through_the_dispatch_on_the_device():
    """fem_solver.parallel = True through Assemble / AssembleExplicit with the real handle: three partitions executed one after the
    other on this GPU, row blocks emitted by fl_row_block_emit."""
    from florence_b200 import assembly
    from test_gpu_plugin import make_objects
    key = "asm_tet2_n2_LinearElastic"
    g = np.load(os.path.join(GOLD, "golden_assembly.npz"))
    n = g[key + "_T"].shape[0]
    Kref = csr_matrix((g[key + "_K_data"], g[key + "_K_indices"], g[key + "_K_indptr"]), shape=(n, n))
    so, fs, fo, me, mat = make_objects(g, key, "LinearElastic", "mechanics", True)
    so.analysis_type, so.is_mass_computed = "static", True
    so.parallel, so.no_of_cpu_cores, so.is_partitioned = True, 3, False
    K, T, F, M = assembly.Assemble(so, fs, fo, me, mat, g[key + "_points"], np.zeros(me.nnode))
    assert abs(K - Kref).max() <= 1e-10 * abs(Kref).max()
    assert all(b["vals"].is_cuda and b["cols"].is_cuda for b in so.row_block)
    so1, *_ = make_objects(g, key, "LinearElastic", "mechanics", True)
    so1.analysis_type, so1.is_mass_computed, so1.parallel = "static", True, False
    K1, T1, _, _ = assembly.Assemble(so1, fs, fo, me, mat, g[key + "_points"], np.zeros(me.nnode))
    assert abs(K - K1).max() <= 1e-13 * abs(K1).max()
    key2 = "asm_hex2_n2_NeoHookean"
    sp, fs2, fo2, me2, mat2 = make_objects(g, key2, "NeoHookean", "mechanics", True)
    sp.is_mass_computed, sp.parallel, sp.no_of_cpu_cores, sp.is_partitioned = True, True, 2, False
    Tp, _, _ = assembly.AssembleExplicit(sp, fs2, fo2, me2, mat2, g[key2 + "_Eulerx"], np.zeros(me2.nnode))
    assert np.linalg.norm(Tp.ravel() - g[key2 + "_T"]) <= 1e-11 * np.linalg.norm(g[key2 + "_T"])
    assembly.clear_handles()


def test_host_results_are_caller_owned_by_default():
    """The reference returns fresh arrays from every call; so does the plug-in unless reuse_host_buffers(True) is set.  Large results
    are leased pinned buffers recycled by a weakref finaliser: never while a reference or a view exists."""
    from florence_b200 import assembly
    from test_gpu_plugin import make_objects
    key = "asm_hex2_n2_NeoHookean"
    g = np.load(os.path.join(GOLD, "golden_assembly.npz"))
    so, fs, fo, me, mat = make_objects(g, key, "NeoHookean", "mechanics", True)
    so.analysis_type, so.is_mass_computed, so.parallel = "static", True, False
    big = assembly._BIG
    try:
        assembly._BIG = 1 << 10          # treat this small K as a "large" result: exercises the pipelined fresh-array path
        Ks = [assembly.Assemble(so, fs, fo, me, mat, g[key + "_Eulerx"] * (1.0 + 1e-3 * k), np.zeros(me.nnode))[0] for k in range(4)]
        datas = [K.data.copy() for K in Ks]
        Ks.append(assembly.Assemble(so, fs, fo, me, mat, g[key + "_Eulerx"], np.zeros(me.nnode))[0])
        for K, d in zip(Ks, datas):
            assert np.array_equal(K.data, d)           # earlier results are not overwritten by later calls
        assert not np.array_equal(Ks[0].data, Ks[1].data)
        # large results are leased pinned buffers: one may only be written again once the result AND every view of it are gone
        import gc
        view = Ks[0].data[::2]
        keep = view.copy()
        addr = Ks[0].data.__array_interface__["data"][0]
        del Ks, K
        gc.collect()
        later = [assembly.Assemble(so, fs, fo, me, mat, g[key + "_Eulerx"] * (1.0 + 2e-3 * k), np.zeros(me.nnode))[0] for k in range(5)]
        assert all(not np.shares_memory(view, L.data) for L in later) and np.array_equal(view, keep)
        del view, later
        gc.collect()
        again = [assembly.Assemble(so, fs, fo, me, mat, g[key + "_Eulerx"], np.zeros(me.nnode))[0] for _ in range(3)]
        assert any(A.data.__array_interface__["data"][0] == addr for A in again)     # ... and then it IS recycled
        del again
        prev = assembly.reuse_host_buffers(True)
        a = assembly._LowLevelAssembly_Par_(so, fs, fo, me, mat, g[key + "_Eulerx"], np.zeros(me.nnode))[2]
        assembly._LowLevelAssembly_Par_(so, fs, fo, me, mat, g[key + "_Eulerx"] * 1.001, np.zeros(me.nnode))
        c = assembly._LowLevelAssembly_Par_(so, fs, fo, me, mat, g[key + "_Eulerx"] * 1.002, np.zeros(me.nnode))[2]
        assert np.shares_memory(a, c)                   # the opt-in ring hands out views of its two pinned buffers
        assembly.reuse_host_buffers(prev)
    finally:
        assembly._BIG = big
        assembly.reuse_host_buffers(False)
        assembly.clear_handles()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_explicit_run_under_torchrun():
    script = os.path.join(os.path.dirname(__file__), "multigpu_check.py")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", script], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTIGPU_CHECK_OK" in res.stdout
