"""CPU: host logic of the reference-facing layer -- Assemble / LowLevelAssembly / install(),
the parallel launchers and the partitioner -- with the device handle replaced by a checker-backed double (tests/_fake_handle.py).
What is tested here is everything between the Florence objects and the C-ABI call: name lookup, argument unpacking, COO / CSR
wrapping, return shapes, error behaviour, partition bookkeeping.  The same entry points run on the real handle in
tests/test_gpu_plugin.py (-m gpu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.sparse import csr_matrix

from _fake_handle import FakeHandle, install_fake
from test_gpu_plugin import make_objects

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _golden(key):
    g = np.load(os.path.join(GOLD, "golden_assembly.npz"))
    n = g[key + "_T"].shape[0]
    Kref = csr_matrix((g[key + "_K_data"], g[key + "_K_indices"], g[key + "_K_indptr"]), shape=(n, n))
    Eulerx = g[key + "_Eulerx"] if int(g[key + "_update"]) else g[key + "_points"]
    Eulerp = g[key + "_Eulerp"] if key + "_Eulerp" in g.files else np.zeros(g[key + "_points"].shape[0])
    return g, Kref, Eulerx, Eulerp


@pytest.mark.parametrize("key,fields", [("asm_hex1_n2_NeoHookean", "mechanics"), ("asm_tet2_n2_LinearElastic", "mechanics"),
                                        ("asm_hex1_n2_IsotropicElectroMechanics_101", "electro_mechanics")])
def test_assemble_and_low_level_assembly(monkeypatch, key, fields):
    """Assembly.py:25-95: Assemble -> LowLevelAssembly -> _LowLevelAssembly_ -> stamped wrapper, both scatter modes."""
    assembly = install_fake(monkeypatch)
    g, Kref, Eulerx, Eulerp = _golden(key)
    matname = key.split("_", 3)[3]
    n = Kref.shape[0]
    for recompute in (True, False):
        so, fs, fo, me, mat = make_objects(g, key, matname, fields, recompute)
        so.analysis_type, so.is_mass_computed, so.parallel = "static", False, False
        if not recompute:
            idx, iptr = assembly.ComputeSparsityPattern(me, fo.nvar, True, function_space=fs)
            so.indices, so.indptr = idx, iptr
        K, T, F, M = assembly.Assemble(so, fs, fo, me, mat, Eulerx, Eulerp)
        assert isinstance(K, csr_matrix) and K.shape == (n, n) and K.dtype == np.float64
        assert T.shape == (n, 1) and F == [] and M == []
        assert so.assembly_time >= 0.0
        assert abs(K - Kref).max() <= 1e-10 * abs(Kref).max()
        # dynamic analysis: the first call also returns the (lumped) mass and sets the flag (Assembly.py:51-73)
        so.analysis_type, so.mass_type = "dynamic", "lumped"
        K2, T2, F2, M2 = assembly.LowLevelAssembly(so, fs, fo, me, mat, Eulerx, Eulerp)
        assert so.is_mass_computed is True and np.asarray(M2).shape == (n, 1)
        # row-sum lumping (negative at the vertices of quadratic simplices, as in the reference); the total mass is rho * volume
        assert abs(np.asarray(M2).reshape(-1, fo.nvar)[:, 0].sum() - 1100.0 * 0.96) <= 1e-9 * 1100.0
        K3, T3, F3, M3 = assembly.LowLevelAssembly(so, fs, fo, me, mat, Eulerx, Eulerp)
        assert M3 == []
        assert abs(K3 - K).max() == 0.0


def test_dispatch_errors_and_laplacian_route(monkeypatch):
    assembly = install_fake(monkeypatch)
    g, Kref, Eulerx, Eulerp = _golden("asm_hex1_n2_NeoHookean")
    so, fs, fo, me, mat = make_objects(g, "asm_hex1_n2_NeoHookean", "NeoHookean", "mechanics", True)
    so.analysis_type, so.is_mass_computed, so.parallel = "static", True, False
    # unknown material: the reference's NotImplementedError (_LowLevelAssembly_.py:58-60)
    other = type("SomeOtherMaterial", (object,), {})()
    with pytest.raises(NotImplementedError):
        assembly.Assemble(so, fs, fo, me, other, Eulerx, Eulerp)
    # material that opts out of the low-level dispatcher: RuntimeError (Assembly.py:41-42, :667-668)
    mat.has_low_level_dispatcher = False
    with pytest.raises(RuntimeError):
        assembly.LowLevelAssembly(so, fs, fo, me, mat, Eulerx, Eulerp)
    with pytest.raises(RuntimeError):
        assembly.AssembleExplicit(so, fs, fo, me, mat, Eulerx, Eulerp)
    # electrostatics goes to the Laplacian wrapper and returns (K, T[:,None], None, None) (Assembly.py:44-47)
    gl = np.load(os.path.join(GOLD, "golden_laplacian.npz"))
    key = "lap_hex2_n2"
    fsl, fol, mel, sol, matl = (type("O", (object,), {})() for _ in range(5))
    fsl.Jm, fsl.AllGauss, fsl.Bases = gl[key + "_Jm"], gl[key + "_AllGauss"], None
    fol.ndim, fol.nvar, fol.fields = 3, 1, "electrostatics"
    mel.points, mel.elements = gl[key + "_points"], gl[key + "_elements"].astype(np.uint64)
    mel.nelem, mel.nnode = mel.elements.shape[0], mel.points.shape[0]
    mel.ChangeType = lambda: None
    mel.GetNumberOfNodes = lambda: mel.nnode
    sol.recompute_sparsity_pattern, sol.squeeze_sparsity_pattern = True, False
    matl.e = gl[key + "_e"]
    K, T, F, M = assembly.Assemble(sol, fsl, fol, mel, matl, None, None)
    nn = mel.nnode
    Kl = csr_matrix((gl[key + "_K_data"], gl[key + "_K_indices"], gl[key + "_K_indptr"]), shape=(nn, nn))
    assert F is None and M is None and T.shape == (nn, 1) and not T.any()
    assert abs(K - Kl).max() <= 1e-10 * abs(Kl).max()


def test_assemble_explicit_returns_mass_on_first_call(monkeypatch):
    """Assembly.py:664-715."""
    assembly = install_fake(monkeypatch)
    g, Kref, Eulerx, Eulerp = _golden("asm_hex2_n2_NeoHookean")
    so, fs, fo, me, mat = make_objects(g, "asm_hex2_n2_NeoHookean", "NeoHookean", "mechanics", True)
    so.is_mass_computed, so.mass_type, so.parallel = False, "lumped", False
    T, F, M = assembly.AssembleExplicit(so, fs, fo, me, mat, Eulerx, Eulerp)
    n = g["asm_hex2_n2_NeoHookean_T"].shape[0]
    assert T.shape == (n, 1) and so.is_mass_computed is True and np.asarray(M).shape == (n, 1)
    assert np.linalg.norm(T.ravel() - g["asm_hex2_n2_NeoHookean_T"]) <= 1e-11 * np.linalg.norm(g["asm_hex2_n2_NeoHookean_T"])
    T2, F2, M2 = assembly.AssembleExplicit(so, fs, fo, me, mat, Eulerx, Eulerp)
    assert M2 == [] and np.array_equal(T2, T)


@pytest.mark.parametrize("n_parts,order", [(3, "sfc"), (2, None)])
def test_parallel_branch_serial_execution(monkeypatch, n_parts, order):
    """fem_solver.parallel = True (Assembly.py:79-82, :672-674) in one process: the partitions are executed one after the other,
    the owned row blocks tile the global matrix, the explicit partial forces sum to the global force."""
    assembly = install_fake(monkeypatch)
    from florence_b200 import parallel
    key = "asm_tet2_n2_LinearElastic"
    g, Kref, Eulerx, Eulerp = _golden(key)
    so, fs, fo, me, mat = make_objects(g, key, "LinearElastic", "mechanics", True)
    so.analysis_type, so.is_mass_computed = "static", True
    so.parallel, so.no_of_cpu_cores, so.is_partitioned = True, n_parts, False
    parallel.PartitionMeshForParallelFEM(so, me, n_parts, fo.nvar, order=order)
    # bookkeeping of FEMSolver.py:1630-1656
    assert len(so.pmesh) == n_parts and so.is_partitioned
    allel = np.sort(np.concatenate(so.pelement_indices))
    assert np.array_equal(allel, np.arange(me.nelem))                          # every element in exactly one block
    for r in range(n_parts):
        assert np.array_equal(so.partitioned_maps[r].reshape(-1, 3)[:, 0], 3 * so.pnode_indices[r])
    K, T, F, M = assembly.LowLevelAssembly(so, fs, fo, me, mat, Eulerx, Eulerp)
    assert abs(K - Kref).max() <= 1e-10 * abs(Kref).max()
    blocks = so.row_block
    assert sum(b["rows"].numel() for b in blocks) == K.shape[0]
    assert blocks[-1]["nnz_offset"] + int(blocks[-1]["indptr"][-1]) == blocks[0]["total_nnz"] == K.nnz
    so2, *_ = make_objects(g, key, "LinearElastic", "mechanics", True)
    so2.analysis_type, so2.is_mass_computed, so2.parallel = "static", True, False
    K1, T1, _, _ = assembly.LowLevelAssembly(so2, fs, fo, me, mat, Eulerx, Eulerp)
    assert abs(K - K1).max() <= 1e-13 * abs(K1).max() and np.abs(T - T1).max() <= 1e-13 * max(np.abs(T1).max(), 1e-300)
    # explicit
    so.requires_geometry_update = True
    key2 = "asm_hex2_n2_NeoHookean"
    g2, _, Ex2, Ep2 = _golden(key2)
    sp, fs2, fo2, me2, mat2 = make_objects(g2, key2, "NeoHookean", "mechanics", True)
    sp.is_mass_computed, sp.parallel, sp.no_of_cpu_cores, sp.is_partitioned = True, True, n_parts, False
    Tp, _, _ = assembly.AssembleExplicit(sp, fs2, fo2, me2, mat2, Ex2, Ep2)
    assert np.linalg.norm(Tp.ravel() - g2[key2 + "_T"]) <= 1e-11 * np.linalg.norm(g2[key2 + "_T"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/Florence"), reason="the reference checkout only exists in the build container")
def test_install_with_the_real_florence_objects(monkeypatch):
    """install(Florence) + the reference's own Assemble with REAL Mesh / FunctionSpace / FEMSolver / material objects, up to the
    native call (the handle is the checker-backed double): proves the duck-typing assumptions of SURVEY.md 8b on the package
    itself.  The result must equal what the reference's Python path (optimise=False) assembles."""
    import sys
    import warnings
    sys.path.insert(0, GOLD)
    import _load_reference
    warnings.simplefilter("ignore")
    Fl = _load_reference.load()
    from Florence import Mesh, FEMSolver, DisplacementFormulation, DisplacementPotentialFormulation, AssembleForm
    import Florence as F
    assembly = install_fake(monkeypatch)
    target = assembly.install(Fl)
    asm_mod = sys.modules["Florence.FiniteElements.Assembly.Assembly"]
    assert asm_mod._LowLevelAssembly_ is assembly._LowLevelAssembly_ and target.has_low_level_dispatcher
    assert asm_mod.ImplicitParallelLauncher is assembly.ImplicitParallelLauncher
    rng = np.random.default_rng(5)
    for etype, p, matname in (("hex", 2, "NeoHookean"), ("tet", 2, "MooneyRivlin"), ("hex", 1, "IsotropicElectroMechanics_108")):
        mesh = Mesh()
        mesh.Parallelepiped(upper_right_front_point=(1.0, 0.8, 1.2), nx=2, ny=2, nz=2, element_type=etype)
        if etype == "tet":
            form0 = DisplacementFormulation(mesh)
            Jm0 = form0.function_spaces[0].Jm
            det = np.array([np.linalg.det(Jm0[:, :, 0] @ mesh.points[e]) for e in mesh.elements])
            mesh.elements[det < 0] = mesh.elements[det < 0][:, [1, 0, 2, 3]]
            mesh.GetBoundaryFacesTet(); mesh.GetBoundaryEdgesTet()
        if p > 1:
            mesh.GetHighOrderMesh(p=p)
        electro = matname.startswith("Isotropic")
        kw = dict(mu=4e5, lamb=2e6) if matname == "NeoHookean" else dict(mu1=2.4e5, mu2=1.6e5, lamb=2e6)
        if electro:
            kw["eps_2"] = 4.0 * 8.8541e-12
        material = getattr(F, matname)(3, rho=1100.0, **kw)
        form = DisplacementPotentialFormulation(mesh) if electro else DisplacementFormulation(mesh)
        Eulerx = mesh.points + 0.01 * rng.uniform(-1, 1, mesh.points.shape)
        Eulerp = 9e3 * mesh.points[:, 2] + 10 * rng.uniform(-1, 1, mesh.points.shape[0]) if electro else np.zeros(mesh.points.shape[0])
        # reference python path
        material.has_low_level_dispatcher = False
        s0 = FEMSolver(analysis_nature="nonlinear", optimise=False, recompute_sparsity_pattern=True)
        K0, T0 = AssembleForm(form, mesh, material, s0, Eulerx=Eulerx.copy(), Eulerp=Eulerp.copy())
        # reference dispatch -> this back end (real objects all the way to the handle)
        material.has_low_level_dispatcher = True
        s1 = FEMSolver(analysis_nature="nonlinear", optimise=True, recompute_sparsity_pattern=True)
        s1.has_low_level_dispatcher = True
        s1.requires_geometry_update = True
        s1.is_mass_computed = True
        out = asm_mod.Assemble(s1, form.function_spaces[0], form, mesh, material, Eulerx.copy(), Eulerp.copy())
        K1, T1 = out[0], out[1]
        n = K0.shape[0]
        nvar = form.nvar
        mech = np.arange(n) % nvar < 3
        for ra in (mech, ~mech):
            for ca in (mech, ~mech):
                if ra.any() and ca.any():
                    A, B = K1.tocsr()[ra][:, ca], K0.tocsr()[ra][:, ca]
                    assert abs(A - B).max() <= 1e-10 * abs(B).max(), (etype, matname)
        assert np.linalg.norm((T1.ravel() - T0.ravel())[mech]) <= 1e-11 * np.linalg.norm(T0.ravel()[mech])
        assembly._handle_cache.clear()


# ---------------------------------------------------------------------------------------------------- world_size 2 / 3 (gloo)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launcher_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from florence_b200 import assembly, parallel
        assembly.AssemblyHandle = FakeHandle
        assembly._to_host = lambda t, tag, defer=False, prepared=None: t.numpy().copy()
        assembly._to_host_many = lambda items, prepared=None: tuple(t.numpy().copy() for t, _ in items)
        key = "asm_tet2_n2_LinearElastic"
        g, Kref, Eulerx, Eulerp = _golden(key)
        so, fs, fo, me, mat = make_objects(g, key, "LinearElastic", "mechanics", True)
        so.analysis_type, so.is_mass_computed = "static", True
        so.parallel, so.no_of_cpu_cores, so.is_partitioned = True, world, False
        K, T, F, M = assembly.LowLevelAssembly(so, fs, fo, me, mat, Eulerx, Eulerp)
        b = so.row_block
        errK = float(abs(K - Kref).max() / abs(Kref).max())
        ok_offsets = (b["total_nnz"] == K.nnz and b["total_rows"] == K.shape[0])
        key2 = "asm_hex2_n2_NeoHookean"
        g2, _, Ex2, Ep2 = _golden(key2)
        sp, fs2, fo2, me2, mat2 = make_objects(g2, key2, "NeoHookean", "mechanics", True)
        sp.is_mass_computed, sp.parallel, sp.no_of_cpu_cores, sp.is_partitioned = True, True, world, False
        Tp, _, _ = assembly.AssembleExplicit(sp, fs2, fo2, me2, mat2, Ex2, Ep2)
        errT = float(np.linalg.norm(Tp.ravel() - g2[key2 + "_T"]) / np.linalg.norm(g2[key2 + "_T"]))
        q.put((rank, errK, errT, bool(ok_offsets), int(b["row_offset"]), int(b["nnz_offset"]), Tp.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_parallel_launchers_one_process_per_partition(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_launcher_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errK, errT, ok_offsets, ro, no, _ in res:
        assert errK <= 1e-10 and errT <= 1e-11 and ok_offsets
    assert res[0][4] == 0 and res[0][5] == 0 and all(res[i][4] < res[i + 1][4] and res[i][5] < res[i + 1][5] for i in range(world - 1))
    # the rank-ordered interface sums make the assembled T bit-identical on every rank
    assert all(r[6] == res[0][6] for r in res)


def test_sfc_order_shrinks_the_interface_of_a_shuffled_mesh():
    """BASELINE north_star: contiguous element blocks cut along a space-filling curve.  A 16^3 hex mesh whose element order has
    been shuffled (what an unstructured mesh file looks like) has almost every node on an interface when cut as it comes
    (Mesh.py:7403); after Morton ordering the interface is within 1.3x of the structured slab cut."""
    from florence_b200 import mesh as flmesh, partition
    pts, els = flmesh.box_hex_mesh(16, 16, 16, p=1)
    P, E = pts.numpy(), els.numpy()
    rng = np.random.default_rng(0)
    Es = E[rng.permutation(E.shape[0])]
    for world in (2, 3, 4, 8):
        slab = partition.interface_node_count(E, world)
        raw = partition.interface_node_count(Es, world)
        sfc = partition.interface_node_count(Es, world, order="sfc", points=P)
        assert raw > 0.9 * P.shape[0]                  # cut as it comes: nearly every node is shared
        assert sfc <= 1.3 * slab, (world, sfc, slab)   # 289/289, 697/614, 561/867, 817/2023 nodes
    # the permutation is a permutation, and partition_mesh(order="sfc") keeps the map back to the caller's element numbers
    perm = partition.sfc_order(P, Es)
    assert np.array_equal(np.sort(perm), np.arange(Es.shape[0]))
    part = partition.partition_mesh(P, Es, 1, 4, order="sfc")
    gl = part.node_map.numpy()
    assert np.array_equal(gl[part.elements.numpy()], Es[part.element_ids.numpy()])


def test_leased_result_buffers_bookkeeping(monkeypatch):
    """Host-result ownership (assembly._lease / _leased_array): at most _LEASE_MAX buffers per size are out, a buffer only goes
    back to the free list when the array AND every view of it are gone, and it is then handed out again.  Page-locked allocation is
    replaced by ordinary memory so that the logic runs without a GPU."""
    import gc
    import torch
    from florence_b200 import assembly
    real_empty = torch.empty
    monkeypatch.setattr(assembly.torch, "empty", lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "pin_memory"}))
    assembly._lease_free.clear(); assembly._lease_out.clear()
    t = torch.arange(1000, dtype=torch.float64)
    held = []
    for _ in range(assembly._LEASE_MAX):
        buf, key = assembly._lease(t)
        assert buf is not None
        buf.copy_(t)
        held.append(assembly._leased_array(buf, key))
    assert assembly._lease(t)[0] is None                    # the limit: further results are pageable arrays
    assert len({a.__array_interface__["data"][0] for a in held}) == assembly._LEASE_MAX
    view = held[0][10:20]
    addr = held[0].__array_interface__["data"][0]
    del held[0]
    gc.collect()
    assert assembly._lease(t)[0] is None                    # a view still refers to the first buffer
    assert np.array_equal(view, np.arange(10, 20, dtype=np.float64))
    del view
    gc.collect()
    buf, key = assembly._lease(t)                           # ... now it is free again, and it is the same memory
    assert buf is not None and buf.data_ptr() == addr
    assembly._lease_return(key, buf)
    del held
    gc.collect()
    assert assembly._lease_out[key] == 0 and len(assembly._lease_free[key]) == assembly._LEASE_MAX
    assembly.release_host_buffers()
    assert not assembly._lease_free
