"""Parallel launchers of the assembly dispatch, one process per GPU.

Drop-ins (same names, argument lists and return values) for
    FEMSolver.PartitionMeshForParallelFEM      Florence/Solver/FEMSolver.py:1630-1656
    ImplicitParallelLauncher                   Florence/FiniteElements/Assembly/Assembly.py:879-1050
    ExplicitParallelLauncher                   Florence/FiniteElements/Assembly/Assembly.py:1126-1358
which the reference reaches through `fem_solver.parallel` (Assembly.py:79-82, :672-676).  The reference partitions the mesh
into `no_of_cpu_cores` contiguous element blocks (Mesh.Partition, Mesh.py:7395-7447), pickles every sub-mesh to a worker
process on each call and sums the returned triplets / tractions on the parent.  Here the SAME script runs once per GPU
(torchrun; rank r <-> partition r), every rank keeps its partition resident on its device, and

  implicit : each rank assembles its block plus the halo elements that complete the rows of the nodes it owns and emits the CSR
             ROW BLOCK it owns, with global column numbers, on the device (fl_row_block_build / fl_row_block_emit); the position
             of the block in the global matrix comes from an all_gather of the row / nnz counts and an exclusive scan.  No K value
             ever crosses a link.  `fem_solver.parallel_gather = True` (default for small problems and for callers that hand K to
             a host solver, as the reference does) additionally all_gathers the blocks into one scipy CSR on every rank.
  explicit : each rank evaluates its block's internal forces and the interface partial sums are exchanged (partition.py); the
             launcher returns T for the whole mesh (all_gather of the owned parts) because its caller indexes T globally.

With a single process (no process group) and `no_of_cpu_cores = n > 1` the n partitions are executed one after the other on
the one visible GPU -- the reference's commented-out SERIAL variant (Assembly.py:929-944) -- which is also how the `-m gpu`
tests exercise these entry points.
"""
import numpy as np
import torch
import torch.distributed as dist
from scipy.sparse import csr_matrix

from . import partition


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def PartitionMeshForParallelFEM(fem_solver, mesh, n, nvar, order="sfc"):
    """FEMSolver.py:1630-1656.  Stores on fem_solver, per partition this process executes: the explicit Partition (block +
    interface lists), the implicit RowPartition (block + halo, owned rows) and the reference's bookkeeping lists
    (pmesh / pelement_indices / pnode_indices / partitioned_maps, entries of other ranks are None).  order="sfc" cuts the blocks
    after ordering the elements along a space-filling curve (BASELINE north_star); order=None is the reference's cut."""
    if getattr(fem_solver, "is_partitioned", False):
        return
    rank, world = _world()
    if world > 1 and n != world:
        raise ValueError("no_of_cpu_cores (%d) must equal the number of ranks (%d): one partition per GPU" % (n, world))
    mine = [rank] if world > 1 else list(range(n))
    pts = np.asarray(mesh.points)
    els = np.asarray(mesh.elements).astype(np.int64)
    if els.shape[0] < n:
        raise ValueError("Could not partition the mesh correctly for parallel processing")
    fem_solver.partitions = {r: partition.partition_mesh(pts, els, r, n, order=order) for r in mine}
    fem_solver.row_partitions = {r: partition.row_partition(pts, els, r, n, order=order) for r in mine}
    map_facilitator = np.arange(pts.shape[0] * nvar, dtype=np.int32).reshape(pts.shape[0], nvar)
    fem_solver.pmesh = [None] * n
    fem_solver.pelement_indices = [None] * n
    fem_solver.pnode_indices = [None] * n
    fem_solver.partitioned_maps = [None] * n
    for r in mine:
        part = fem_solver.partitions[r]
        fem_solver.pmesh[r] = part
        fem_solver.pelement_indices[r] = part.element_ids.numpy()
        fem_solver.pnode_indices[r] = part.node_map.numpy()
        fem_solver.partitioned_maps[r] = map_facilitator[part.node_map.numpy(), :].ravel()
    fem_solver.is_partitioned = True
    fem_solver.no_of_cpu_cores = n


class _LocalMesh(object):
    """The duck-typed mesh the assembler wrappers read (points, elements, nelem, ChangeType)."""

    def __init__(self, points, elements):
        self.points = points.numpy() if isinstance(points, torch.Tensor) else np.asarray(points)
        e = elements.numpy() if isinstance(elements, torch.Tensor) else np.asarray(elements)
        self.elements = np.ascontiguousarray(e.astype(np.uint64))
        self.nelem = self.elements.shape[0]
        self.nnode = self.points.shape[0]

    def ChangeType(self):
        pass

    def GetNumberOfNodes(self):
        return self.nnode


def _device_row_block(fem_solver, function_space, formulation, lmesh, material, Eulerx_l, Eulerp_l, rp):
    """Default local assembler of ImplicitParallelLauncher: CSR assembly of the rank's sub-mesh and emission of the owned row
    block on the device.  Returns (indptr_block, cols_global, vals, T_owned) as numpy arrays plus the device tensors."""
    from . import assembly
    nvar = formulation.nvar
    h = assembly.get_handle(lmesh, function_space)
    mat = assembly._material_struct(material)
    form = 1 if formulation.fields == "electro_mechanics" else 0
    x, p = assembly._state_to_device(h, Eulerx_l, Eulerp_l if form else None)
    h.build_pattern(nvar)
    V, T = h.assemble_implicit(x, p, mat, form, bool(fem_solver.requires_geometry_update), mode="csr")
    indptr, cols, vals = h.row_block(nvar, V, rp.owned_local, rp.node_map)
    own = torch.as_tensor(rp.owned_local).long().to(T.device)
    T_owned = T.view(-1, nvar)[own]
    return indptr, cols, vals, T_owned


def ImplicitParallelLauncher(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp, local_assembler=None):
    """Assembly.py:879-1050: returns (stiffness, T.ravel(), F, mass).

    Every rank owns a row block; fem_solver.row_block = dict(rows=global row ids, indptr, cols, vals, row_offset, nnz_offset,
    total_rows, total_nnz) keeps it on the device for a distributed solver.  The returned `stiffness` is the full matrix when
    fem_solver.parallel_gather is true (default), else a (n x n) scipy CSR holding only this rank's rows.
    `local_assembler(fem_solver, function_space, formulation, local_mesh, material, Eulerx_local, Eulerp_local, row_partition)`
    -> (indptr_block, cols_global, vals, T_owned) replaces the device path (the CPU tests inject the checker there)."""
    nvar = formulation.nvar
    nnode = mesh.points.shape[0]
    n = nvar * nnode
    if not getattr(fem_solver, "is_partitioned", False):
        PartitionMeshForParallelFEM(fem_solver, mesh, fem_solver.no_of_cpu_cores, nvar)
    rank, world = _world()
    asm = local_assembler or _device_row_block
    Eulerx = np.asarray(Eulerx)
    Eulerp = None if Eulerp is None else np.asarray(Eulerp)
    blocks = []
    for r in sorted(fem_solver.row_partitions):
        rp = fem_solver.row_partitions[r]
        gl = rp.node_map.numpy()
        lmesh = getattr(rp, "_lmesh", None)
        if lmesh is None:
            lmesh = rp._lmesh = _LocalMesh(rp.points, rp.elements)
        indptr, cols, vals, T_owned = asm(fem_solver, function_space, formulation, lmesh, material, Eulerx[gl],
                                          None if Eulerp is None else Eulerp[gl], rp)
        rows = rp.global_rows(nvar)
        blocks.append(dict(rank=r, rows=rows, indptr=indptr, cols=cols, vals=vals, T_owned=T_owned))
    if world > 1:
        b = blocks[0]
        dev = b["vals"].device if isinstance(b["vals"], torch.Tensor) else "cpu"
        ro, no, tr, tn = partition.global_row_offsets(b["rows"].numel(), int(b["indptr"][-1]), device=dev)
        b.update(row_offset=ro, nnz_offset=no, total_rows=tr, total_nnz=tn)
        fem_solver.row_block = b
    else:
        ro = no = 0
        for b in blocks:       # serial execution of all partitions: offsets by the same exclusive scan
            b.update(row_offset=ro, nnz_offset=no)
            ro += b["rows"].numel()
            no += int(b["indptr"][-1])
        for b in blocks:
            b.update(total_rows=ro, total_nnz=no)
        fem_solver.row_block = blocks[0] if len(blocks) == 1 else blocks

    def host(t):
        return t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)

    T = np.zeros((nnode, nvar), np.float64)
    gather = getattr(fem_solver, "parallel_gather", True)
    pieces = [(host(b["rows"]), host(b["indptr"]), host(b["cols"]), host(b["vals"]), host(b["T_owned"])) for b in blocks]
    if world > 1 and gather:
        allp = [None] * world
        dist.all_gather_object(allp, pieces[0])
        pieces = allp
    I, J, V = [], [], []
    for rows, indptr, cols, vals, T_owned in pieces:
        counts = np.diff(indptr)
        I.append(np.repeat(rows, counts))
        J.append(cols)
        V.append(vals)
        T.reshape(-1)[rows] = np.asarray(T_owned).reshape(-1)
    I, J, V = np.concatenate(I), np.concatenate(J), np.concatenate(V)
    # rows are disjoint between ranks and (row, col) pairs unique within a block: no summation happens here, only placement
    stiffness = csr_matrix((V, (I, J)), shape=(n, n), dtype=np.float64)
    F, mass = [], []
    return stiffness, T.ravel(), F, mass


def _device_partial_forces(function_space, formulation, lmesh, material, Eulerx_l, Eulerp_l, part):
    from . import assembly
    return assembly._LowLevelAssemblyExplicit_DF_DPF_(function_space, formulation, lmesh, material, Eulerx_l, Eulerp_l, device_out=True)


def ExplicitParallelLauncher(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp, local_assembler=None):
    """Assembly.py:1126-1358: returns T_all.ravel() (nnode*nvar) -- `T_all[pnodes,:] += T_p` over the partitions (:1352-1354), the
    interface nodes summed in ascending partition order on every rank."""
    nvar = formulation.nvar
    nnode = mesh.points.shape[0]
    if not getattr(fem_solver, "is_partitioned", False):
        PartitionMeshForParallelFEM(fem_solver, mesh, fem_solver.no_of_cpu_cores, nvar)
    rank, world = _world()
    asm = local_assembler or _device_partial_forces
    Eulerx = np.asarray(Eulerx)
    Eulerp = None if Eulerp is None else np.asarray(Eulerp)
    partial = {}
    for r in sorted(fem_solver.partitions):
        part = fem_solver.partitions[r]
        gl = part.node_map.numpy()
        lmesh = getattr(part, "_lmesh", None)
        if lmesh is None:
            lmesh = part._lmesh = _LocalMesh(part.points, part.elements)
        Tl = asm(function_space, formulation, lmesh, material, Eulerx[gl], None if Eulerp is None else Eulerp[gl], part)
        partial[r] = (gl, Tl, lmesh)
    T_all = np.zeros((nnode, nvar), np.float64)
    if world > 1:
        part = fem_solver.partitions[rank]
        gl, Tl, lmesh = partial[rank]
        on_device = isinstance(Tl, torch.Tensor) and Tl.is_cuda
        ex = getattr(part, "_exchange", None)
        if ex is None or ex.nvar != nvar:
            if on_device:      # NCCL: the partial sums never leave the device before they are complete
                from . import assembly
                ex = partition.InterfaceExchange(part, nvar, Tl.device, handle=assembly.get_handle(lmesh, function_space))
            else:
                ex = partition.InterfaceExchange(part, nvar, "cpu")
            part._exchange = ex
        Tt = Tl.reshape(-1) if on_device else torch.from_numpy(np.ascontiguousarray(np.asarray(Tl, dtype=np.float64).reshape(-1)))
        ex(Tt)                                   # rank-ordered interface sums; now exact on every local node
        allp = [None] * world
        dist.all_gather_object(allp, (gl, Tt.cpu().numpy().reshape(-1, nvar)))
        for g, t in allp:
            T_all[g] = t
    else:
        for r in sorted(partial):                # ascending partition order, as the parent of the reference sums
            gl, Tl, _ = partial[r]
            T_all[gl] += (Tl.cpu().numpy() if isinstance(Tl, torch.Tensor) else np.asarray(Tl)).reshape(-1, nvar)
    return T_all.ravel()
