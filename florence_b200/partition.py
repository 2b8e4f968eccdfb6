"""Element partitioning across GPUs (one process per GPU) and the interface exchange of the explicit path.

Replaces the reference's process-pool / MPI launchers (Florence/FiniteElements/Assembly/Assembly.py:879-1050, :1126-1358),
FEMSolver.PartitionMeshForParallelFEM (Florence/Solver/FEMSolver.py:1630-1656) and Mesh.Partition
(Florence/MeshGeneration/Mesh.py:7395-7447): the same contiguous element blocks (np.array_split), cut after the elements have
been ordered along a space-filling curve (BASELINE north_star), but each rank keeps its block on its own GPU, integrates its own
nodes, and only the partial internal forces of INTERFACE nodes cross NVLink -- the reference broadcasts the whole Eulerx and
reduces the whole T every step (Assembly.py:1153-1163).

Exchange: the partial T of the shared nodes is packed (fl_pack_nodes / fl_gather_pack_nodes), swapped with torch.distributed
P2P ops (NCCL over NVLink on GPUs, gloo in the CPU tests) and summed by fl_sum_ordered: for every shared node the contributions
of ALL ranks that hold it, the rank's own included, are added in ascending rank order, so the replicated interface dofs are
bit-identical on every sharer however many ranks meet at a node.
"""
import numpy as np
import torch
import torch.distributed as dist


class Partition(object):
    """Local view of one rank: local mesh, local->global node map, per-neighbour interface lists (local node ids)."""

    def __init__(self, rank, world, points, elements, node_map, neighbours, element_ids=None, n_interface_elements=None):
        self.rank, self.world = rank, world
        self.points, self.elements = points, elements
        self.node_map = node_map            # global node id of each local node (ascending)
        self.neighbours = neighbours        # {other_rank: int32 tensor of local node ids, ordered by global id}
        self.element_ids = element_ids      # global element id of each local element (None: the contiguous block itself)
        # local elements [0, n_interface_elements) touch an interface node (interface_first()); None = not ordered that way
        self.n_interface_elements = n_interface_elements

    def interface_nodes(self):
        """Sorted unique local ids of the nodes shared with any other rank."""
        if not self.neighbours:
            return torch.zeros(0, dtype=torch.int32, device=self.elements.device if isinstance(self.elements, torch.Tensor) else "cpu")
        return torch.unique(torch.cat([v.reshape(-1) for v in self.neighbours.values()])).to(torch.int32)

    def interface_first(self):
        """Reorder the local elements so that those touching an interface node come first (stable within both groups): the
        explicit step then evaluates them first, starts the exchange, and evaluates the interior elements while the messages
        are in flight (SURVEY.md 8e "boundary elements first")."""
        els = self.elements
        nn = self.points.shape[0]
        flag = torch.zeros(nn, dtype=torch.bool, device=els.device)
        iface = self.interface_nodes().to(els.device).long()
        flag[iface] = True
        touches = flag[els.long()].any(dim=1)
        order = torch.argsort((~touches).to(torch.int8), stable=True)
        self.elements = els[order].contiguous()
        if self.element_ids is not None:
            self.element_ids = self.element_ids[order.to(self.element_ids.device)]
        else:
            self.element_order = order          # position in the rank's contiguous block
        self.n_interface_elements = int(touches.sum().item())
        return self


def element_blocks(nelem, world):
    """Contiguous blocks, identical to np.array_split(np.arange(nelem), world) (Mesh.py:7403)."""
    sizes = [nelem // world + (1 if r < nelem % world else 0) for r in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(starts[r]), int(starts[r + 1])) for r in range(world)]


# ------------------------------------------------------------------------------------------------ space-filling-curve order
def _spread3(x):
    x = x & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_keys(points, elements):
    """Host twin of the device key kernel (csrc/fl_pattern.cu morton_keys_kernel): 21 bits per axis of the element centroid
    relative to the bounding box of the nodes, every operation individually rounded in the same order."""
    pts = np.asarray(points, dtype=np.float64)
    els = np.asarray(elements).astype(np.int64)
    nelem, npe = els.shape
    d = pts.shape[1]
    key = np.zeros(nelem, dtype=np.uint64)
    lo, hi = pts.min(0), pts.max(0)
    for k in range(d):
        c = np.zeros(nelem)
        for a in range(npe):                      # sequential sum, as the kernel
            c = c + pts[els[:, a], k]
        c = c / float(npe)
        span = hi[k] - lo[k]
        u = (c - lo[k]) / span if span > 0 else np.zeros(nelem)
        q = np.clip((u * 2097152.0).astype(np.int64), 0, 2097151).astype(np.uint64)
        key |= _spread3(q) << np.uint64(k)
    return key


def sfc_order(points, elements):
    """Permutation that sorts the elements along a Morton curve through their centroids (stable).  CUDA tensors are ordered on
    their device (fl_sfc_order: key kernel + radix sort); anything else takes the numpy twin, which gives the same permutation."""
    if isinstance(elements, torch.Tensor) and elements.is_cuda:
        from . import backend
        return backend.sfc_order(points, elements)
    pts = points.cpu().numpy() if isinstance(points, torch.Tensor) else np.asarray(points)
    els = elements.cpu().numpy() if isinstance(elements, torch.Tensor) else np.asarray(elements)
    return np.argsort(morton_keys(pts, els), kind="stable")


def partition_mesh(points, elements, rank, world, order=None):
    """Generic partitioner (any mesh): block of elements -> localised mesh + interface lists.
    points (nnode x d) and elements (nelem x npe) are the GLOBAL arrays.  order="sfc" sorts the elements along a space-filling
    curve before the contiguous cut (Partition.element_ids then maps local elements back to the caller's numbering); None keeps
    the reference's cut of the incoming order (Mesh.py:7403).  numpy / CPU tensors take the host path; CUDA tensors are
    localised on the device (torch.unique / searchsorted / isin), which is what makes Mesh.Partition's GetLocalisedMesh
    (Mesh.py:7406-7409) usable at 10^7-10^8 elements."""
    if order not in (None, "sfc"):
        raise ValueError("order must be None or 'sfc'")
    if isinstance(elements, torch.Tensor) and elements.is_cuda:
        return _partition_mesh_device(points, elements, rank, world, order)
    pts = np.asarray(points)
    els = np.asarray(elements).astype(np.int64)
    perm = None
    if order == "sfc":
        perm = sfc_order(pts, els)
        els = els[perm]
    blocks = element_blocks(els.shape[0], world)
    node_sets = [np.unique(els[b0:b1]) for (b0, b1) in blocks]        # pnode_indices of Mesh.Partition (Mesh.py:7406-7409)
    mine = node_sets[rank]
    b0, b1 = blocks[rank]
    local = np.searchsorted(mine, els[b0:b1])
    neighbours = {}
    for r in range(world):
        if r == rank:
            continue
        shared = np.intersect1d(mine, node_sets[r], assume_unique=True)
        if shared.size:
            neighbours[r] = torch.as_tensor(np.searchsorted(mine, shared).astype(np.int32))
    eids = torch.as_tensor(perm[b0:b1].astype(np.int64)) if perm is not None else torch.arange(b0, b1, dtype=torch.int64)
    return Partition(rank, world, torch.as_tensor(pts[mine]), torch.as_tensor(local), torch.as_tensor(mine), neighbours, element_ids=eids)


def _partition_mesh_device(points, elements, rank, world, order=None):
    """partition_mesh on the device the mesh lives on: same result, tensors stay on that device."""
    els = elements.long()
    pts = points if isinstance(points, torch.Tensor) else torch.as_tensor(np.asarray(points))
    pts = pts.to(els.device)
    perm = None
    if order == "sfc":
        perm = sfc_order(pts, els)
        els = els[perm]
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
    eids = perm[b0:b1] if perm is not None else torch.arange(b0, b1, dtype=torch.int64, device=els.device)
    return Partition(rank, world, pts[mine], local, mine, neighbours, element_ids=eids)


def interface_node_count(elements, world, order=None, points=None):
    """Number of distinct nodes shared by at least two of the `world` contiguous element blocks (a measure of the cut)."""
    els = np.asarray(elements).astype(np.int64)
    if order == "sfc":
        els = els[sfc_order(points, els)]
    count = np.zeros(int(els.max()) + 1, dtype=np.int32)
    for b0, b1 in element_blocks(els.shape[0], world):
        count[np.unique(els[b0:b1])] += 1
    return int((count > 1).sum())


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
    """Sums a nodal vector over the ranks that share each interface node, every sharer adding the same terms in the same
    (ascending rank) order.

    Layout: U = sorted union of the local ids shared with any neighbour; `all` = [own partials at U | received partials of
    neighbour r0 | r1 | ...] (one contiguous buffer, the receive buffers are views of it); (ptr, idx) list, per node of U, the
    offsets in `all` of its contributions sorted by rank.  With a device `handle` the pack / ordered-sum / scatter steps are the
    library's kernels (fl_pack_nodes, fl_gather_pack_nodes, fl_sum_ordered, fl_scatter_nodes); without one (CPU tests) torch ops
    do the same arithmetic in the same order.
    """

    def __init__(self, partition, nvar, device, pack=None, unpack_add=None, handle=None):
        self.part, self.nvar = partition, nvar
        self.device = torch.device(device)
        self.handle = handle
        self.ranks = sorted(partition.neighbours)
        ids = {r: partition.neighbours[r].to("cpu").numpy().astype(np.int64).reshape(-1) for r in self.ranks}
        U = np.unique(np.concatenate([ids[r] for r in self.ranks])) if self.ranks else np.zeros(0, np.int64)
        nU = U.shape[0]
        self.n_interface = nU
        pos = {r: np.searchsorted(U, ids[r]) for r in self.ranks}
        off, offsets = nU * nvar, {}
        for r in self.ranks:
            offsets[r] = off
            off += ids[r].shape[0] * nvar
        # contributions (slot, rank, offset), sorted by slot then rank
        slot = [np.arange(nU)] + [pos[r] for r in self.ranks]
        rk = [np.full(nU, partition.rank)] + [np.full(ids[r].shape[0], r) for r in self.ranks]
        of = [np.arange(nU) * nvar] + [offsets[r] + np.arange(ids[r].shape[0]) * nvar for r in self.ranks]
        slot, rk, of = np.concatenate(slot), np.concatenate(rk), np.concatenate(of)
        o = np.lexsort((rk, slot))
        ptr = np.zeros(nU + 1, dtype=np.int64)
        np.cumsum(np.bincount(slot, minlength=nU), out=ptr[1:])
        dev = self.device
        self.U = torch.as_tensor(U.astype(np.int32), device=dev)
        self.pos = {r: torch.as_tensor(pos[r].astype(np.int32), device=dev) for r in self.ranks}
        self.ptr = torch.as_tensor(ptr, device=dev)
        self.idx = torch.as_tensor(of[o].astype(np.int64), device=dev)
        self.all = torch.zeros(max(off, 1), dtype=torch.float64, device=dev)
        self.own = self.all[:nU * nvar]
        self.recv = {r: self.all[offsets[r]:offsets[r] + ids[r].shape[0] * nvar] for r in self.ranks}
        self.send = {r: torch.empty(ids[r].shape[0] * nvar, dtype=torch.float64, device=dev) for r in self.ranks}
        self.T_iface = torch.zeros(max(nU * nvar, 1), dtype=torch.float64, device=dev)
        nn = partition.points.shape[0]
        self.iface_slot = torch.full((nn,), -1, dtype=torch.int32, device=dev)
        if nU:
            self.iface_slot[self.U.long()] = torch.arange(nU, dtype=torch.int32, device=dev)
        self._works = []

    def bytes_per_exchange(self):
        return sum(b.numel() * 8 for b in self.send.values())

    # ---- building blocks (device kernels, or torch on the CPU)
    def _ptr(self, t):
        import ctypes as C
        return C.c_void_p(t.data_ptr())

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _pack(self, T, ids, buf):
        if self.handle is not None:
            from ._lib import check
            check(self.handle.lib.fl_pack_nodes(self._ptr(T), self._ptr(ids), ids.numel(), self.nvar, self._ptr(buf), self._stream()))
        else:
            buf.copy_(T.view(-1, self.nvar)[ids.long()].reshape(-1))

    def _sum_ordered(self):
        n = self.n_interface
        if n == 0:
            return
        if self.handle is not None:
            from ._lib import check
            check(self.handle.lib.fl_sum_ordered(self._ptr(self.all), self._ptr(self.ptr), self._ptr(self.idx), n, self.nvar,
                                                 self._ptr(self.T_iface), self._stream()))
            return
        out = torch.zeros(n, self.nvar, dtype=torch.float64)
        cnt = self.ptr[1:] - self.ptr[:-1]
        comp = torch.arange(self.nvar)
        for j in range(int(cnt.max().item())):
            m = cnt > j
            src = self.idx[self.ptr[:-1][m] + j]
            out[m] = out[m] + self.all[src[:, None] + comp[None, :]]
        self.T_iface[:n * self.nvar].copy_(out.reshape(-1))

    def _scatter(self, T):
        n = self.n_interface
        if n == 0:
            return
        if self.handle is not None:
            from ._lib import check
            check(self.handle.lib.fl_scatter_nodes(self._ptr(T), self._ptr(self.U), n, self.nvar, self._ptr(self.T_iface), self._stream()))
        else:
            T.view(-1, self.nvar)[self.U.long()] = self.T_iface[:n * self.nvar].view(-1, self.nvar)

    # ---- split-phase exchange (with overlap=True the explicit step evaluates the interior elements between start and finish)
    def start(self, own_filled=False, T=None):
        """Post the sends / receives of the partial sums in `own` (filled by the caller when own_filled, else packed from T)."""
        if not self.ranks:
            return
        if not own_filled:
            self._pack(T, self.U, self.own)
        for r in self.ranks:
            self._pack(self.own, self.pos[r], self.send[r])
        ops = []
        for r in self.ranks:
            ops.append(dist.P2POp(dist.isend, self.send[r], r))
            ops.append(dist.P2POp(dist.irecv, self.recv[r], r))
        self._works = dist.batch_isend_irecv(ops)

    def finish(self):
        """Wait for the messages and form the rank-ordered sums in T_iface (one slot per node of U)."""
        for w in self._works:
            w.wait()
        self._works = []
        self._sum_ordered()

    def __call__(self, T):
        """In place: T[shared nodes] = sum over all sharers, ascending rank order."""
        if not self.ranks:
            return T
        self.start(T=T)
        self.finish()
        self._scatter(T)
        return T


def device_pack_functions(handle):
    """Kept for callers of the round-1 API: InterfaceExchange(..., handle=handle) now runs the library's kernels itself."""
    return None, None


# ------------------------------------------------------------------------------------------------ implicit: owned CSR row blocks
class RowPartition(Partition):
    """Partition for implicit assembly: the local mesh holds every element that touches a node owned by this rank (the rank's
    block plus halo elements), so the CSR rows of owned nodes are complete without any communication
    ("each rank emits the CSR row block it owns")."""

    def __init__(self, rank, world, points, elements, node_map, owned_local, n_halo_elements, element_ids=None):
        Partition.__init__(self, rank, world, points, elements, node_map, {}, element_ids=element_ids)
        self.owned_local = owned_local            # local node ids owned by this rank (ascending global id)
        self.n_halo_elements = n_halo_elements

    def owned_rows(self, V, indices, indptr, nvar):
        """Host restatement of the row-block emission (what AssemblyHandle.row_block does on the device; kept as its checker):
        slice the locally assembled CSR (V, indices, indptr over local dofs) down to the owned rows and translate the column
        indices to GLOBAL dof numbers.  Returns (global_row_ids, indptr_block, global_cols, values)."""
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

    def global_rows(self, nvar):
        """Global dof numbers of the owned rows, in block order."""
        gl = torch.as_tensor(self.node_map).long()
        own = torch.as_tensor(self.owned_local).long().to(gl.device)
        return (gl[own][:, None] * nvar + torch.arange(nvar, device=gl.device)[None, :]).reshape(-1)


def row_partition(points, elements, rank, world, order=None):
    """Node ownership = lowest rank whose contiguous element block (np.array_split, Mesh.py:7403) contains the node;
    local elements = every element with at least one owned node.  order="sfc": blocks are cut after Morton ordering."""
    pts = points.cpu().numpy() if isinstance(points, torch.Tensor) else np.asarray(points)
    els = (elements.cpu().numpy() if isinstance(elements, torch.Tensor) else np.asarray(elements)).astype(np.int64)
    perm = np.arange(els.shape[0])
    if order == "sfc":
        perm = sfc_order(pts, els)
        els = els[perm]
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
                        torch.as_tensor(owned_local), n_halo, element_ids=torch.as_tensor(perm[touches]))


def global_row_offsets(n_rows_local, nnz_local, device="cpu"):
    """Position of this rank's row block in the global CSR: (row offset, nnz offset, total rows, total nnz) from an all_gather of
    the per-rank counts followed by an exclusive scan.  Without an initialised process group: offsets 0."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, 0, int(n_rows_local), int(nnz_local)
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([int(n_rows_local), int(nnz_local)], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    counts = torch.stack(allc).cpu().numpy()
    before = counts[:rank].sum(axis=0) if rank > 0 else np.zeros(2, np.int64)
    total = counts.sum(axis=0)
    return int(before[0]), int(before[1]), int(total[0]), int(total[1])
