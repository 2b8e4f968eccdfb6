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
        q.put((rank, float(errT), float(errM), bool(ok_blocks), int(sum(v.numel() for v in part.neighbours.values()))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("generic_tet", 2), ("slab_hex", 2), ("slab_hex", 3)])
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
    for rank, errT, errM, ok_blocks, nshared in res:
        assert errT < 1e-13 and errM < 1e-13, (rank, errT, errM)
        assert ok_blocks
        assert nshared > 0
