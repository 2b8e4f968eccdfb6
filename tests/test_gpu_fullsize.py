"""GPU: BASELINE.json's full-size configs, checked through size-independent properties (the oracle cannot run these sizes in
seconds): rigid-body null space, symmetry, equilibrium of internal forces, translation invariance, COO/CSR checksum,
run-to-run bit reproducibility, and agreement with the oracle on a random sample of elements."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_config2_tet10_linear_elastic_1M_elements():
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    dev = torch.device("cuda:0")
    n = 55
    pts, els = flmesh.box_tet_mesh(n, n, n, p=2, device=dev)
    assert els.shape == (998250, 10) and pts.shape == (1367631, 3)
    Bases, Jm, AG = flmesh.tables("tet", 2)
    x = flmesh.perturbed_state(pts, 1.0 / n, 1e-3 * n, seed=0)
    mu, lamb = 1.0e5, 1.5e5
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases, device=dev)
    mat = backend.make_material(10, 0.0, mu=mu, lamb=lamb)
    indices, indptr = h.sparsity_pattern(3)
    nrow = 3 * pts.shape[0]
    assert indptr[-1].item() == indices.numel() and indices.numel() == h.nnz[3]
    # pattern: sorted, duplicate-free rows, diagonal present
    rowlen = (indptr[1:] - indptr[:-1]).long()
    assert int(rowlen.min()) > 0
    d = indices[1:] - indices[:-1]
    starts = torch.zeros(indices.numel(), dtype=torch.bool, device=dev)
    starts[indptr[1:-1].long()] = True
    assert bool((d[~starts[1:]] > 0).all())
    V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
    V2, T2 = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
    assert torch.equal(V, V2) and torch.equal(T, T2)            # deterministic reduction
    # the default path stores K_e along the Morton curve and reduces it with the register-resident gather (fl_stream.cu): at full
    # size it must equal, bit for bit, the element-order path with the shared-memory row-buffer reduction, and the register gather
    # in element order
    for o3, o4 in ((0, 0), (1, 0)):
        h.set_option(3, o3); h.set_option(4, o4)
        V2, T2 = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
        assert torch.equal(V, V2) and torch.equal(T, T2), (o3, o4)
    h.set_option(3, 2); h.set_option(4, 1)
    del V2, T2
    K = torch.sparse_csr_tensor(indptr.long(), indices.long(), V, size=(nrow, nrow))
    scale = float(V.abs().max())
    # rigid translations are in the null space of K (row sums of each displacement component vanish)
    for c in range(3):
        r = torch.zeros(nrow, dtype=torch.float64, device=dev)
        r[c::3] = 1.0
        assert float((K @ r).abs().max()) <= 1e-9 * scale
    # symmetry: y^T K z == z^T K y for random vectors
    g = torch.Generator(device=dev); g.manual_seed(3)
    y = torch.rand(nrow, dtype=torch.float64, device=dev, generator=g)
    z = torch.rand(nrow, dtype=torch.float64, device=dev, generator=g)
    a, b = float(y @ (K @ z)), float(z @ (K @ y))
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    # internal forces are self-equilibrated
    assert float(T.view(-1, 3).sum(0).abs().max()) <= 1e-9 * float(T.abs().max()) * 1e3
    # COO mode carries the same matrix: checksum of checksums, and T is identical
    I, J, Vc, Tc = h.assemble_implicit(x, None, mat, 0, True, mode="coo")
    assert torch.equal(Tc, T)
    assert abs(float(Vc.sum()) - float(V.sum())) <= 1e-9 * scale
    w = torch.rand(nrow, dtype=torch.float64, device=dev, generator=g)
    lhs = float((w[I.long()] * Vc * z[J.long()]).sum())
    rhs = float(w @ (K @ z))
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs))
    # ... and row by row: K z from the CSR values against the same product accumulated from the 898 M COO triplets
    Kz = K @ z
    Kz_coo = torch.zeros(nrow, dtype=torch.float64, device=dev).index_add_(0, I.long(), Vc * z[J.long()])
    assert float((Kz - Kz_coo).abs().max()) <= 1e-10 * float(Kz.abs().max())
    del Kz, Kz_coo
    # a random sample of element matrices against the oracle (COO layout is element-major)
    rng = np.random.default_rng(0)
    sample = np.sort(rng.choice(els.shape[0], 64, replace=False))
    P, X = pts.cpu().numpy(), x.cpu().numpy()
    E = els[torch.as_tensor(sample, device=dev)].cpu().numpy()
    Io, Jo, Vo, To = orc.assemble_implicit(P, E, X, None, Jm, AG, 3, 6, 1, orc.params(mu=mu, lamb=lamb), 10, mode="coo")
    Vs = Vc.view(-1, 900)[torch.as_tensor(sample, device=dev)].reshape(-1).cpu().numpy()
    assert np.abs(Vs - Vo).max() <= 1e-10 * np.abs(Vo).max()
    h.close()


def test_config3_hex27_neohookean_explicit_8M_elements():
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    dev = torch.device("cuda:0")
    n = 200
    pts, els = flmesh.box_hex_mesh(n, n, n, p=2, device=dev)
    assert els.shape == (8000000, 27)
    Bases, Jm, AG = flmesh.tables("hex", 2)
    x = flmesh.perturbed_state(pts, 0.5 / n, 0.02, seed=0)
    mu, lamb = 4.0e5, 2.0e6
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases, device=dev)
    mat = backend.make_material(1, 1100.0, mu=mu, lamb=lamb)
    T = h.assemble_explicit(x, None, mat, 0)
    assert torch.equal(T, h.assemble_explicit(x, None, mat, 0))
    Tmax = float(T.abs().max())
    assert float(T.view(-1, 3).sum(0).abs().max()) <= 1e-8 * Tmax * 1e3          # self-equilibrated
    shift = torch.tensor([0.25, -0.5, 0.125], dtype=torch.float64, device=dev)   # exactly representable translation
    Ts = h.assemble_explicit(x + shift, None, mat, 0)
    assert float((Ts - T).abs().max()) <= 1e-9 * Tmax
    T0 = h.assemble_explicit(pts, None, mat, 0)                                   # undeformed: F = I -> sigma = 0
    assert float(T0.abs().max()) <= 1e-9 * Tmax
    # scalar kernel and tensor-core kernel agree
    h.set_option(0, 0)
    Tsc = h.assemble_explicit(x, None, mat, 0)
    h.set_option(0, 1)
    assert float((Tsc - T).abs().max()) <= 1e-11 * Tmax
    # nodes whose whole element patch lies inside a sampled slab: compare with the oracle there
    k0 = 3
    sel = torch.arange(k0 * n * n, (k0 + 3) * n * n, device=dev)[:: 997][:40]
    E = els[sel].cpu().numpy()
    nodes = np.unique(E)
    remap = -np.ones(pts.shape[0], dtype=np.int64); remap[nodes] = np.arange(nodes.size)
    P, X = pts[torch.as_tensor(nodes, device=dev)].cpu().numpy(), x[torch.as_tensor(nodes, device=dev)].cpu().numpy()
    # per-element tractions: assemble each sampled element alone on both sides
    hs = backend.AssemblyHandle(P, remap[E], Jm, AG, Bases, device=dev)
    Tl = hs.assemble_explicit(X, None, mat, 0).cpu().numpy()
    To = orc.assemble_explicit(P, remap[E], X, None, Jm, AG, 3, orc.params(mu=mu, lamb=lamb), 1)
    assert np.linalg.norm(Tl - To) <= 1e-11 * np.linalg.norm(To)
    hs.close(); h.close()


def test_config1_poisson_p4_hex_6cubed():
    """BASELINE.json configs[0]: simple_laplace -- Poisson on a 6^3 hex mesh, p=4 (hex125), IdealDielectric eps=2.35; the whole
    K against the oracle (the CPU-runnable case of the reference)."""
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    from scipy.sparse import csr_matrix
    pts, els = flmesh.box_hex_mesh(6, 6, 6, p=4)
    assert els.shape == (216, 125) and pts.shape == (15625, 3)
    Bases, Jm, AG = flmesh.tables("hex", 4)
    e = -2.35 * np.eye(3)                                  # the wrapper passes -material.e (.pyx:81)
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases)
    indices, indptr = h.sparsity_pattern(1)
    V = h.assemble_laplacian(e, True, mode="csr").cpu().numpy()
    pat = orc.sparsity_pattern(els.numpy(), pts.shape[0], 1)
    assert np.array_equal(indices.cpu().numpy(), pat[0]) and np.array_equal(indptr.cpu().numpy(), pat[1])
    Vo = orc.assemble_laplacian(pts.numpy(), els.numpy(), Jm, AG, e, True, mode="csr", pattern=pat)
    assert np.abs(V - Vo).max() <= 1e-10 * np.abs(Vo).max()
    K = csr_matrix((V, pat[0], pat[1]), shape=(15625, 15625))
    assert abs(K - K.T).max() == 0.0                       # upper triangle mirrored, as the reference does for a symmetric tensor
    assert np.abs(K @ np.ones(15625)).max() <= 1e-10 * np.abs(V).max()   # constants are in the null space
    h.close()


def test_config5_hex8_mooney_rivlin_explicit_8M_elements():
    from florence_b200 import backend, mesh as flmesh
    dev = torch.device("cuda:0")
    n = 200
    pts, els = flmesh.box_hex_mesh(n, n, n, p=1, device=dev)
    assert els.shape == (8000000, 8)
    Bases, Jm, AG = flmesh.tables("hex", 1)
    x = flmesh.perturbed_state(pts, 1.0 / n, 0.02, seed=0)
    mu1, nu = 1.0e6, 0.495
    lamb = 2.0 * mu1 * nu / (1.0 - 2.0 * nu)
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases, device=dev)
    mat = backend.make_material(2, 8000.0, mu1=mu1, mu2=0.0, lamb=lamb)
    T = h.assemble_explicit(x, None, mat, 0)
    Tmax = float(T.abs().max())
    assert torch.equal(T, h.assemble_explicit(x, None, mat, 0))
    assert float(T.view(-1, 3).sum(0).abs().max()) <= 1e-8 * Tmax * 1e3
    h.set_option(0, 0)
    Ts = h.assemble_explicit(x, None, mat, 0)
    h.set_option(0, 1)
    assert float((Ts - T).abs().max()) <= 1e-11 * Tmax
    # ExplicitMooneyRivlin (stress-only kernel of the car-crash example) gives the same forces
    mat0 = backend.make_material(0, 8000.0, mu1=mu1, mu2=0.0, lamb=lamb)
    assert torch.equal(h.assemble_explicit(x, None, mat0, 0), T)
    # ten device-resident central-difference steps stay finite and move only free dofs
    M = h.assemble_mass(8000.0, 3, "lumped")
    assert abs(float(M.sum()) / 3.0 - 8000.0) <= 1e-9 * 8000.0          # total mass of the unit cube
    fixed = torch.zeros(pts.shape[0] * 3, dtype=torch.uint8, device=dev)
    fixed[: 3 * (n + 1) * (n + 1)] = 1                                   # clamp the z = 0 plane
    Eulerx = x.reshape(-1).clone()
    U0 = torch.zeros_like(T); U00 = torch.zeros_like(T)
    dt = 0.2 * (1.0 / n) / np.sqrt((lamb + 2 * mu1) / 8000.0)
    status = h.explicit_steps(mat, dt, 10, 2, M, None, fixed, None, U0, U00, Eulerx, T)
    assert status == 0 and bool(torch.isfinite(Eulerx).all())
    assert float((Eulerx - pts.reshape(-1))[: 3 * (n + 1) * (n + 1)].abs().max()) == 0.0
    h.close()


def test_config4_hex64_em108_24cubed():
    """Config 4 at its benchmarked size: hex64 IsotropicElectroMechanics_108, 24^3 elements, K (CSR through the dof-pair-plane scratch
    and the wide node-centric reduction, and COO through the element-major write path) + T.  Properties that do not need the oracle
    at this size, plus 32 sampled element matrices against the oracle (_LowLevelAssemblyDPF_.h:45-198)."""
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    dev = torch.device("cuda:0")
    n = 24
    pts, els = flmesh.box_hex_mesh(n, n, n, p=3, device=dev)
    assert els.shape == (13824, 64) and pts.shape == (389017, 3)
    Bases, Jm, AG = flmesh.tables("hex", 3)
    x = flmesh.perturbed_state(pts, 1.0 / (3 * n), 0.02, seed=11)
    g = torch.Generator(device=dev); g.manual_seed(5)
    phi = 9.0e3 * pts[:, 2] + 10.0 * (2.0 * torch.rand(pts.shape[0], dtype=torch.float64, device=dev, generator=g) - 1.0)
    mu = 5.0e4
    prm = dict(mu1=mu, mu2=mu, lamb=2.0 * mu * 0.4 / (1.0 - 0.8), eps_2=4.0 * 8.8541e-12)
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases, device=dev)
    mat = backend.make_material(8, 1200.0, **prm)
    indices, indptr = h.sparsity_pattern(4)
    nrow = 4 * pts.shape[0]
    V, T = h.assemble_implicit(x, phi, mat, 1, True, mode="csr")
    V2, T2 = h.assemble_implicit(x, phi, mat, 1, True, mode="csr")
    assert torch.equal(V, V2) and torch.equal(T, T2)            # deterministic reduction
    K = torch.sparse_csr_tensor(indptr.long(), indices.long(), V, size=(nrow, nrow))
    mech = (torch.arange(nrow, device=dev) % 4) != 3
    # the four blocks live on very different scales: scale every check by the block it touches
    rows_of = indptr.new_zeros(indices.numel(), dtype=torch.int64)
    rows_of[indptr[1:-1].long()] = 1
    rows_of = torch.cumsum(rows_of, 0)
    rm, cm = mech[rows_of], mech[indices.long()]
    s_uu = float(V[rm & cm].abs().max()); s_up = float(V[rm & ~cm].abs().max()); s_pp = float(V[~rm & ~cm].abs().max())
    # null space: a rigid translation produces no force and no charge; a constant potential shift neither (E = -grad phi)
    for c in range(3):
        r = torch.zeros(nrow, dtype=torch.float64, device=dev)
        r[c::4] = 1.0
        y = K @ r
        assert float(y[mech].abs().max()) <= 1e-9 * s_uu and float(y[~mech].abs().max()) <= 1e-9 * s_up
    r = torch.zeros(nrow, dtype=torch.float64, device=dev)
    r[3::4] = 1.0
    y = K @ r
    assert float(y[mech].abs().max()) <= 1e-9 * s_up and float(y[~mech].abs().max()) <= 1e-9 * s_pp
    # symmetry, block by block: y^T K z == z^T K y with y, z supported on one field each
    ym = torch.rand(nrow, dtype=torch.float64, device=dev, generator=g)
    zm = torch.rand(nrow, dtype=torch.float64, device=dev, generator=g)
    for ya, za in ((mech, mech), (mech, ~mech), (~mech, ~mech)):
        yy, zz = ym * ya, zm * za
        a, b = float(yy @ (K @ zz)), float(zz @ (K @ yy))
        assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
    # COO mode (element-major triplets, different write path and no reduction kernel) carries the same matrix and the same T
    I, J, Vc, Tc = h.assemble_implicit(x, phi, mat, 1, True, mode="coo")
    assert torch.equal(Tc, T)
    for ya, za in ((mech, mech), (mech, ~mech), (~mech, mech), (~mech, ~mech)):
        yy, zz = ym * ya, zm * za
        rhs = float(yy @ (K @ zz))
        lhs = 0.0
        step = 1 << 27
        for k0 in range(0, Vc.numel(), step):
            sl = slice(k0, min(k0 + step, Vc.numel()))
            lhs += float((yy[I[sl].long()] * Vc[sl] * zz[J[sl].long()]).sum())
        assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs))
    # 32 sampled element matrices against the oracle, block-wise
    rng = np.random.default_rng(0)
    sample = np.sort(rng.choice(els.shape[0], 32, replace=False))
    st = torch.as_tensor(sample, device=dev)
    P, X, PH = pts.cpu().numpy(), x.cpu().numpy(), phi.cpu().numpy()
    E = els[st].cpu().numpy()
    Io, Jo, Vo, To = orc.assemble_implicit(P, E, X, PH, Jm, AG, 4, 9, 1, orc.params(**prm), 8, mode="coo")
    Vs = Vc.view(-1, 256, 256)[st].cpu().numpy()
    Vo = Vo.reshape(-1, 256, 256)
    m = np.arange(256) % 4 != 3
    for ra in (m, ~m):
        for ca in (m, ~m):
            A, B = Vs[:, ra][:, :, ca], Vo[:, ra][:, :, ca]
            assert np.abs(A - B).max() <= 1e-10 * np.abs(B).max()
    del I, J, Vc
    h.close()
