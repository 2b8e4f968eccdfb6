"""Element partitioning across GPUs (one process per GPU) and the interface exchange of the explicit path.

Replaces the reference's process-pool / MPI launchers (Florence/FiniteElements/Assembly/Assembly.py:879-1050, :1126-1358),
FEMSolver.PartitionMeshForParallelFEM (Florence/Solver/FEMSolver.py:1630-1656) and Mesh.Partition
(Florence/MeshGeneration/Mesh.py:7395-7447): the same contiguous element blocks (np.array_split), but each rank keeps its
block on its own GPU, integrates its own nodes, and only the partial internal forces of INTERFACE nodes cross NVLink -- the
reference broadcasts the whole Eulerx and reduces the whole T every step (Assembly.py:1153-1163).

Exchange: for every neighbour rank, the partial T of the shared nodes is packed (fl_pack_nodes), swapped with
torch.distributed P2P ops (NCCL over NVLink on GPUs, gloo in the CPU tests) and added (fl_unpack_add_nodes).  Shared values
are summed in rank order on every owner so the replicated interface dofs stay bit-identical.
"""
import numpy as np
import torch
import torch.distributed as dist


class Partition(object):
    """Local view of one rank: local mesh, local->global node map, per-neighbour interface lists (local node ids)."""

    def __init__(self, rank, world, points, elements, node_map, neighbours):
        self.rank, self.world = rank, world
        self.points, self.elements = points, elements
        self.node_map = node_map            # global node id of each local node (ascending)
        self.neighbours = neighbours        # {other_rank: int32 tensor of local node ids, ordered by global id}


def element_blocks(nelem, world):
    """Contiguous blocks, identical to np.array_split(np.arange(nelem), world) (Mesh.py:7403)."""
    sizes = [nelem // world + (1 if r < nelem % world else 0) for r in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(starts[r]), int(starts[r + 1])) for r in range(world)]


def partition_mesh(points, elements, rank, world):
    """Generic partitioner (any mesh): block of elements -> localised mesh + interface lists.
    points (nnode x d) and elements (nelem x npe) are the GLOBAL arrays.  numpy / CPU tensors take the host path; CUDA tensors are
    localised on the device (torch.unique / searchsorted / isin), which is what makes Mesh.Partition's GetLocalisedMesh
    (Mesh.py:7406-7409) usable at 10^7-10^8 elements."""
    if isinstance(elements, torch.Tensor) and elements.is_cuda:
        return _partition_mesh_device(points, elements, rank, world)
    pts = np.asarray(points)
    els = np.asarray(elements).astype(np.int64)
    blocks = element_blocks(els.shape[0], world)
    node_sets = [np.unique(els[b0:b1]) for (b0, b1) in blocks]        # pnode_indices of Mesh.Partition (Mesh.py:7406-7409)
    mine = node_sets[rank]
    local = np.searchsorted(mine, els[blocks[rank][0]:blocks[rank][1]])
    neighbours = {}
    for r in range(world):
        if r == rank:
            continue
        shared = np.intersect1d(mine, node_sets[r], assume_unique=True)
        if shared.size:
            neighbours[r] = torch.as_tensor(np.searchsorted(mine, shared).astype(np.int32))
    return Partition(rank, world, torch.as_tensor(pts[mine]), torch.as_tensor(local), torch.as_tensor(mine), neighbours)


def _partition_mesh_device(points, elements, rank, world):
    """partition_mesh on the device the mesh lives on: same result, tensors stay on that device."""
    els = elements.long()
    pts = points if isinstance(points, torch.Tensor) else torch.as_tensor(np.asarray(points))
    pts = pts.to(els.device)
    blocks = element_blocks(els.shape[0], world)
    b0, b1 = blocks[rank]
    mine = torch.unique(els[b0:b1])                                   # sorted
    local = torch.searchsorted(mine, els[b0:b1].reshape(-1)).reshape(b1 - b0, -1)
    neighbours = {}
    for r in range(world):
        if r == rank:
            continue
        other = torch.unique(els[blocks[r][0]:blocks[r][1]])
        shared = mine[torch.isin(mine, other, assume_unique=True)]
        if shared.numel():
            neighbours[r] = torch.searchsorted(mine, shared).to(torch.int32)
    return Partition(rank, world, pts[mine], local, mine, neighbours)


def slab_partition_hex(nx, ny, nz_per_rank, p, rank, world, lengths_per_rank=(1.0, 1.0, 1.0), device="cpu"):
    """Structured weak-scaling partition: `world` boxes of nx*ny*nz_per_rank order-p hexahedra stacked along z (SURVEY.md 8d
    config 5).  Built directly on the rank's device without ever forming the global mesh; interface = one node plane."""
    from . import mesh as flmesh
    L = (lengths_per_rank[0], lengths_per_rank[1], lengths_per_rank[2] * world)
    pts, els = flmesh.box_hex_mesh(nx, ny, nz_per_rank, p=p, lengths=L, device=device, z_offset_elems=rank * nz_per_rank,
                                   nz_total=nz_per_rank * world)
    NX, NY, NZ = p * nx + 1, p * ny + 1, p * nz_per_rank + 1
    plane = NX * NY
    neighbours = {}
    if rank > 0:
        neighbours[rank - 1] = torch.arange(0, plane, dtype=torch.int32, device=device)
    if rank < world - 1:
        neighbours[rank + 1] = torch.arange(plane * (NZ - 1), plane * NZ, dtype=torch.int32, device=device)
    node0 = rank * p * nz_per_rank * plane
    node_map = torch.arange(node0, node0 + plane * NZ, dtype=torch.int64, device=device)
    return Partition(rank, world, pts, els, node_map, neighbours)


class InterfaceExchange(object):
    """Sums a nodal vector over the ranks that share each interface node."""

    def __init__(self, partition, nvar, device, pack=None, unpack_add=None):
        self.part, self.nvar = partition, nvar
        self.device = torch.device(device)
        self.ids = {r: ids.to(self.device) for r, ids in partition.neighbours.items()}
        self.send = {r: torch.empty(ids.numel() * nvar, dtype=torch.float64, device=self.device) for r, ids in self.ids.items()}
        self.recv = {r: torch.empty_like(b) for r, b in self.send.items()}
        self._pack, self._unpack_add = pack, unpack_add

    def bytes_per_exchange(self):
        return sum(b.numel() * 8 for b in self.send.values())

    def _pack_nodes(self, T, ids, buf):
        if self._pack is not None:
            self._pack(T, ids, buf)
        else:
            buf.copy_(T.view(-1, self.nvar)[ids.long()].reshape(-1))

    def _unpack_add_nodes(self, T, ids, buf):
        if self._unpack_add is not None:
            self._unpack_add(T, ids, buf)
        else:
            T.view(-1, self.nvar)[ids.long()] += buf.view(-1, self.nvar)

    def __call__(self, T):
        """In place: T[shared nodes] += partial sums of the neighbours."""
        if not self.ids:
            return T
        ops = []
        for r in sorted(self.ids):
            self._pack_nodes(T, self.ids[r], self.send[r])
        for r in sorted(self.ids):
            ops.append(dist.P2POp(dist.isend, self.send[r], r))
            ops.append(dist.P2POp(dist.irecv, self.recv[r], r))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        # add in ascending rank order; with one neighbour per shared node (slabs) own+other == other+own bitwise
        for r in sorted(self.ids):
            self._unpack_add_nodes(T, self.ids[r], self.recv[r])
        return T


def device_pack_functions(handle):
    """pack / unpack_add closures running the library's kernels on the handle's device."""
    import ctypes as C
    from ._lib import check
    lib = handle.lib

    def _s():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def pack(T, ids, buf):
        check(lib.fl_pack_nodes(C.c_void_p(T.data_ptr()), C.c_void_p(ids.data_ptr()), ids.numel(), buf.numel() // ids.numel(),
                                C.c_void_p(buf.data_ptr()), _s()))

    def unpack_add(T, ids, buf):
        check(lib.fl_unpack_add_nodes(C.c_void_p(T.data_ptr()), C.c_void_p(ids.data_ptr()), ids.numel(), buf.numel() // ids.numel(),
                                      C.c_void_p(buf.data_ptr()), _s()))
    return pack, unpack_add


# ------------------------------------------------------------------------------------------------ implicit: owned CSR row blocks
class RowPartition(Partition):
    """Partition for implicit assembly: the local mesh holds every element that touches a node owned by this rank (the rank's
    block plus halo elements), so the CSR rows of owned nodes are complete without any communication
    ("each rank emits the CSR row block it owns")."""

    def __init__(self, rank, world, points, elements, node_map, owned_local, n_halo_elements):
        Partition.__init__(self, rank, world, points, elements, node_map, {})
        self.owned_local = owned_local            # local node ids owned by this rank (ascending global id)
        self.n_halo_elements = n_halo_elements

    def owned_rows(self, V, indices, indptr, nvar):
        """Slice the locally assembled CSR (V, indices, indptr over local dofs) down to the owned rows and translate the
        column indices to GLOBAL dof numbers.  Returns (global_row_ids, indptr_block, global_cols, values)."""
        V, indices, indptr = np.asarray(V), np.asarray(indices), np.asarray(indptr)
        gl = np.asarray(self.node_map)
        rows_local = (np.asarray(self.owned_local)[:, None] * nvar + np.arange(nvar)[None, :]).ravel()
        starts, ends = indptr[rows_local], indptr[rows_local + 1]
        counts = ends - starts
        take = np.concatenate([np.arange(s, e) for s, e in zip(starts, ends)]) if rows_local.size else np.zeros(0, np.int64)
        cols_local = indices[take]
        cols_global = gl[cols_local // nvar] * nvar + cols_local % nvar
        rows_global = (gl[np.asarray(self.owned_local)][:, None] * nvar + np.arange(nvar)[None, :]).ravel()
        return rows_global, np.concatenate([[0], np.cumsum(counts)]), cols_global, V[take]


def row_partition(points, elements, rank, world):
    """Node ownership = lowest rank whose contiguous element block (np.array_split, Mesh.py:7403) contains the node;
    local elements = every element with at least one owned node."""
    pts = np.asarray(points)
    els = np.asarray(elements).astype(np.int64)
    blocks = element_blocks(els.shape[0], world)
    owner = np.full(pts.shape[0], world, dtype=np.int64)
    for r in range(world - 1, -1, -1):
        owner[np.unique(els[blocks[r][0]:blocks[r][1]])] = r
    touches = (owner[els] == rank).any(axis=1)
    local_els = els[touches]
    b0, b1 = blocks[rank]
    n_halo = int(touches.sum() - touches[b0:b1].sum())
    nodes = np.unique(local_els)
    local = np.searchsorted(nodes, local_els)
    owned_local = np.nonzero(owner[nodes] == rank)[0]
    return RowPartition(rank, world, torch.as_tensor(pts[nodes]), torch.as_tensor(local), torch.as_tensor(nodes),
                        torch.as_tensor(owned_local), n_halo)
