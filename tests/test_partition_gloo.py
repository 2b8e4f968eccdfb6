"""N>1 host logic on CPU: element-block partitioning + interface exchange (gloo, world_size 2 and 3).

The per-rank element computation is done by the CPU oracle here (tests may use it); on the GPU box the same Partition /
InterfaceExchange objects drive the CUDA kernels (tests/test_gpu_parity.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from florence_b200 import mesh as flmesh, partition
from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prm = orc.params(mu=4e5, lamb=2e6)
        if kind == "generic_tet":
            pts, els = flmesh.box_tet_mesh(3, 2, 4, p=2)
            Bases, Jm, AG = flmesh.tables("tet", 2)
            part = partition.partition_mesh(pts.numpy(), els.numpy(), rank, world)
        else:
            nx, ny, nzr, p = 3, 2, 2, 2
            Bases, Jm, AG = flmesh.tables("hex", p)
            part = partition.slab_partition_hex(nx, ny, nzr, p, rank, world, lengths_per_rank=(1.0, 0.8, 0.5))
            pts, els = flmesh.box_hex_mesh(nx, ny, nzr * world, p=p, lengths=(1.0, 0.8, 0.5 * world))
        P = pts.numpy()
        X = P + 0.01 * np.sin(7.0 * P + 0.3)                       # deterministic global state
        Tg = orc.assemble_explicit(P, els.numpy(), X, None, Jm, AG, 3, prm, 1).reshape(-1, 3)
        Mg = orc.assemble_mass(P, els.numpy(), Bases, Jm, AG, 3, 1100.0, "lumped").reshape(-1, 3)
        gl = part.node_map.numpy()
        # local geometry must be the global one restricted to the rank's nodes
        assert np.allclose(part.points.numpy(), P[gl], atol=1e-14)
        xl = X[gl]
        Tl = orc.assemble_explicit(part.points.numpy(), part.elements.numpy(), xl, None, Jm, AG, 3, prm, 1)
        Ml = orc.assemble_mass(part.points.numpy(), part.elements.numpy(), Bases, Jm, AG, 3, 1100.0, "lumped")
        ex = partition.InterfaceExchange(part, 3, "cpu")
        Tt, Mt = torch.from_numpy(Tl.copy()), torch.from_numpy(Ml.copy())
        ex(Tt); ex(Mt)
        errT = np.abs(Tt.numpy().reshape(-1, 3) - Tg[gl]).max() / np.abs(Tg).max()
        errM = np.abs(Mt.numpy().reshape(-1, 3) - Mg[gl]).max() / np.abs(Mg).max()
        # element blocks are the reference's contiguous np.array_split blocks
        blocks = partition.element_blocks(els.shape[0], world)
        ref = np.array_split(np.arange(els.shape[0]), world)
        ok_blocks = all(b[0] == r[0] and b[1] == r[-1] + 1 for b, r in zip(blocks, ref))
        Tn = Tt.numpy().reshape(-1, 3)
        shared = {int(gl[i]): Tn[i].tobytes() for i in ex.U.numpy()}
        q.put((rank, float(errT), float(errM), bool(ok_blocks), int(sum(v.numel() for v in part.neighbours.values())), shared))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("generic_tet", 2), ("generic_tet", 3), ("generic_tet", 4), ("slab_hex", 2), ("slab_hex", 3)])
def test_partition_and_interface_exchange(kind, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    copies, most = {}, 0
    for rank, errT, errM, ok_blocks, nshared, shared in res:
        assert errT < 1e-13 and errM < 1e-13, (rank, errT, errM)
        assert ok_blocks
        assert nshared > 0
        for node, bits in shared.items():
            copies.setdefault(node, []).append(bits)
    # every rank that holds a shared node computed the same sum BIT FOR BIT (contributions are added in ascending rank order on
    # all sharers), also where three or more ranks meet
    for node, lst in copies.items():
        assert len(lst) >= 2 and all(b == lst[0] for b in lst), node
        most = max(most, len(lst))
    if kind == "generic_tet" and world == 3:      # 144 tets in 3 blocks of 48: the cuts fall inside hex layers, 3 ranks meet at nodes
        assert most >= 3


def _row_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scipy.sparse import csr_matrix
        pts, els = flmesh.box_tet_mesh(3, 3, 4, p=2)
        Bases, Jm, AG = flmesh.tables("tet", 2)
        P, E = pts.numpy(), els.numpy()
        X = P + 0.01 * np.sin(5.0 * P + 0.1)
        prm = orc.params(mu=1e5, lamb=1.5e5)
        part = partition.row_partition(P, E, rank, world)
        gl = part.node_map.numpy()
        lp, le = part.points.numpy(), part.elements.numpy()
        pat = orc.sparsity_pattern(le, lp.shape[0], 3)
        V, T = orc.assemble_implicit(lp, le, X[gl], None, Jm, AG, 3, 6, 1, prm, 10, mode="csr", pattern=pat)
        rows, iptr, cols, vals = part.owned_rows(V, pat[0], pat[1], 3)
        gathered = [None] * world
        dist.all_gather_object(gathered, (rows, iptr, cols, vals, T.reshape(-1, 3)[part.owned_local.numpy()], gl[part.owned_local.numpy()]))
        if rank == 0:
            n = 3 * P.shape[0]
            patg = orc.sparsity_pattern(E, P.shape[0], 3)
            Vg, Tg = orc.assemble_implicit(P, E, X, None, Jm, AG, 3, 6, 1, prm, 10, mode="csr", pattern=patg)
            Kg = csr_matrix((Vg, patg[0], patg[1]), shape=(n, n))
            allrows = np.concatenate([g[0] for g in gathered])
            ok_cover = np.array_equal(np.sort(allrows), np.arange(n))         # every row owned exactly once
            err = 0.0
            errT = 0.0
            for rws, ip, cl, vl, Tl, own in gathered:
                for k, r in enumerate(rws):
                    ref_cols = Kg.indices[Kg.indptr[r]:Kg.indptr[r + 1]]
                    ref_vals = Kg.data[Kg.indptr[r]:Kg.indptr[r + 1]]
                    c, v = cl[ip[k]:ip[k + 1]], vl[ip[k]:ip[k + 1]]
                    if not np.array_equal(c, ref_cols):
                        err = np.inf
                        break
                    err = max(err, np.abs(v - ref_vals).max())
                errT = max(errT, np.abs(Tl - Tg.reshape(-1, 3)[own]).max())
            q.put((bool(ok_cover), float(err / np.abs(Vg).max()), float(errT / np.abs(Tg).max()), int(part.n_halo_elements)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_owned_csr_row_blocks_need_no_communication(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_row_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok_cover, err, errT, n_halo = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_cover and err < 1e-13 and errT < 1e-13 and n_halo > 0
