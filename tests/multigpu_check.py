"""torchrun script (not collected by pytest): a partitioned explicit run on N GPUs against the same run on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multigpu_check.py

Checks: (1) the N-rank trajectory equals the single-domain trajectory within fp64 rounding, (2) the replicated interface dofs are
bit-identical on the ranks sharing them after every chunk, (3) the overlapped step (interface elements first, interior elements
during the exchange) is bit-identical to the non-overlapped one, (4) a blow-up stops every rank at the same increment,
(5) the implicit row blocks of all ranks tile the global matrix with the offsets from the all_gather scan."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from florence_b200 import backend, mesh as flmesh, partition, time_integrator  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nx, ny, nz, p = 6, 5, 8, 2
    pts, els = flmesh.box_hex_mesh(nx, ny, nz, p=p, lengths=(1.0, 0.8, 1.4))
    B, Jm, AG = flmesh.tables("hex", p)
    mu, lamb, rho = 4.0e5, 2.0e6, 1100.0
    mat = backend.make_material(1, rho, mu=mu, lamb=lamb)
    hx = 1.0 / (p * nx)
    dt = 0.2 * hx / np.sqrt((lamb + 2 * mu) / rho)
    x0 = pts + 0.02 * hx * torch.sin(40.0 * pts + 0.3)
    fixed = torch.zeros(pts.shape[0] * 3, dtype=torch.uint8)
    fixed.view(-1, 3)[pts[:, 2] == 0] = 1
    fext = torch.zeros(pts.shape[0], 3, dtype=torch.float64)
    fext[pts[:, 2] == pts[:, 2].max(), 0] = 5.0
    nsteps = 24

    def run(part, overlap, order_tag):
        hl = backend.AssemblyHandle(part.points, part.elements, Jm, AG, B, device=dev)
        ex = partition.InterfaceExchange(part, 3, dev, handle=hl) if part.world > 1 else None
        integ = time_integrator.ExplicitStructuralDynamicIntegrator(hl, mat, rho=rho, exchange=ex, overlap=overlap)
        gl = part.node_map.to("cpu")
        integ.initialise(part.points, None, fixed.view(-1, 3)[gl].reshape(-1), dt)
        integ.Eulerx.copy_(x0[gl].reshape(-1).to(dev))
        integ.internal_force(integ.Eulerx.view(-1, 3), out=integ.T)
        f = fext[gl].reshape(-1).to(dev)
        snaps = []
        for c0 in range(2, 2 + nsteps, 8):
            st = integ.step(8, c0, f, 0.0, 1.0 / (nsteps + 2))
            assert st == 0, (order_tag, st)
            snaps.append(integ.displacement().cpu().clone())
        return integ, hl, snaps, gl

    # single-domain reference on every rank (small)
    whole = partition.Partition(0, 1, pts, els, torch.arange(pts.shape[0]), {})
    _, h0, ref, _ = run(whole, False, "single")
    scale = max(float(r.abs().max()) for r in ref)
    for order in (None, "sfc"):
        part = partition.partition_mesh(pts.numpy(), els.numpy(), rank, world, order=order).interface_first()
        part.points, part.elements = part.points.to(dev), part.elements.to(dev)
        i1, h1, s_ov, gl = run(part, True, "overlap")
        i2, h2, s_no, _ = run(part, False, "plain")
        for a, b, r in zip(s_ov, s_no, ref):
            assert torch.equal(a, b), "overlapped and plain steps differ"
            err = float((a - r[gl]).abs().max())
            assert err <= 1e-10 * scale, (order, err, scale)
        # interface dofs: identical bits on all sharers
        ex = i1.exchange
        U = ex.U.long().cpu()
        mine = {int(gl[i]): s_ov[-1][i].numpy().tobytes() for i in U}
        allm = [None] * world
        dist.all_gather_object(allm, mine)
        for other in allm:
            for node, bits in other.items():
                if node in mine:
                    assert mine[node] == bits, "interface dof differs between ranks"
        h1.close(); h2.close()
    # blow-up: every rank reports the same status and increment
    part = partition.partition_mesh(pts.numpy(), els.numpy(), rank, world, order="sfc").interface_first()
    hl = backend.AssemblyHandle(part.points, part.elements, Jm, AG, B, device=dev)
    ex = partition.InterfaceExchange(part, 3, dev, handle=hl)
    integ = time_integrator.ExplicitStructuralDynamicIntegrator(hl, mat, rho=rho, exchange=ex)
    gl = part.node_map
    integ.initialise(part.points, None, fixed.view(-1, 3)[gl].reshape(-1), 8.0 * dt)
    integ.Eulerx.copy_(x0[gl].reshape(-1).to(dev))
    integ.internal_force(integ.Eulerx.view(-1, 3), out=integ.T)
    st = integ.step(60, 2, fext[gl].reshape(-1).to(dev), 0.0, 1.0 / 62)
    got = [None] * world
    dist.all_gather_object(got, (st, integ.last_status))
    assert st != 0 and all(g == got[0] for g in got), got
    hl.close()
    # implicit: owned row blocks + offsets
    tp, te = flmesh.box_tet_mesh(4, 4, 5, p=2)
    Bt, Jt, At = flmesh.tables("tet", 2)
    xt = flmesh.perturbed_state(tp, 0.2, 0.02, seed=2)
    mle = backend.make_material(10, 1.0, mu=1e5, lamb=1.5e5)
    rp = partition.row_partition(tp.numpy(), te.numpy(), rank, world, order="sfc")
    hr = backend.AssemblyHandle(rp.points, rp.elements, Jt, At, Bt, device=dev)
    hr.build_pattern(3)
    V, T = hr.assemble_implicit(xt[rp.node_map].to(dev), None, mle, 0, True, mode="csr")
    ip, cols, vals = hr.row_block(3, V, rp.owned_local, rp.node_map)
    ro, no, tr, tn = partition.global_row_offsets(ip.numel() - 1, int(ip[-1]), device=dev)
    hg = backend.AssemblyHandle(tp, te, Jt, At, Bt, device=dev)
    nnz = hg.build_pattern(3)
    Vg, Tg = hg.assemble_implicit(xt.to(dev), None, mle, 0, True, mode="csr")
    ig, pg = hg.sparsity_pattern(3)
    assert tr == 3 * tp.shape[0] and tn == nnz
    rows = rp.global_rows(3).to(dev)
    for k in np.random.default_rng(rank).choice(rows.numel(), 200, replace=False):
        r = int(rows[k]); s, e = int(pg[r]), int(pg[r + 1])
        a, b = int(ip[k]), int(ip[k + 1])
        assert torch.equal(cols[a:b], ig[s:e].long())
        assert float((vals[a:b] - Vg[s:e]).abs().max()) <= 1e-13 * float(Vg.abs().max())
    allo = [None] * world
    dist.all_gather_object(allo, (ro, no, ip.numel() - 1, int(ip[-1])))
    if rank == 0:
        acc_r = acc_n = 0
        for (r0, n0, nr, nn_) in allo:
            assert r0 == acc_r and n0 == acc_n
            acc_r += nr; acc_n += nn_
        print("MULTIGPU_CHECK_OK world=%d interface_nodes=%d" % (world, ex.n_interface))
    hr.close(); hg.close(); h0.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
