"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the reference-generated golden fixtures.

Tolerances (fp64, BASELINE.json north_star): stiffness entries |dK| <= 1e-10 * max|K| (per block for electro-mechanics, whose
blocks live on scales 20 orders of magnitude apart) AND per entry: relative 1e-10 on every entry above 1e-3 * max|K| of its
block, relative 1e-9 on every entry above 1e-6 * max|K| (entries that small are sums of terms of size max|K| that cancel to
six digits, so one rounding of a term is already 1e-10 of the entry; the CPU oracle itself sits at 8e-11 of the reference
there); residual ||dT|| <= 1e-11 ||T||; sparsity pattern and slot maps bit-exact.
"""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ELEC = ["IsotropicElectroMechanics_101", "IsotropicElectroMechanics_105", "IsotropicElectroMechanics_108"]
GOLD = os.path.join(os.path.dirname(__file__), "golden")
GOLD_FILES = ("golden_assembly.npz", "golden_assembly_hi.npz", "golden_assembly_tet3.npz")   # tet3: cubic tetrahedra (tet20, 14 Gauss points)


def _cases():
    out = []
    for f in GOLD_FILES:
        out += [str(s) for s in np.load(os.path.join(GOLD, f))["asm_cases"]]
    return out


def _load(key):
    for f in GOLD_FILES:
        g = np.load(os.path.join(GOLD, f))
        if key + "_points" in g.files:
            break
    names = ("points", "elements", "Eulerx", "Jm", "AllGauss", "Bases", "K_data", "K_indices", "K_indptr", "T", "update", "prm",
             "sp_indices", "sp_indptr")
    c = {n: g[key + "_" + n] for n in names}
    c["Eulerp"] = g[key + "_Eulerp"] if key + "_Eulerp" in g.files else None
    c["sp_dl"] = g[key + "_sp_dl"] if key + "_sp_dl" in g.files else None
    c["sp_dg"] = g[key + "_sp_dg"] if key + "_sp_dg" in g.files else None
    if int(c["update"]) == 0:
        c["Eulerx"] = c["points"]
    return c


def _material(backend, num, prm):
    return backend.make_material(num, 0.0, mu=prm[0], mu1=prm[1], mu2=prm[2], mu3=prm[3], mue=prm[4], lamb=prm[5], eps_1=prm[6],
                                 eps_2=prm[7], eps_3=prm[8], eps_e=prm[9])


def _entrywise(A, B, tol):
    """max-norm bound + per-entry relative bounds (module docstring); A, B sparse with any pattern."""
    D = np.abs((A - B).toarray())
    R = np.abs(B.toarray())
    mx = R.max()
    assert D.max() <= tol * mx
    if tol < 1e-9:
        return          # a tighter-than-parity identity check (e.g. CSR == summed COO): the max-norm bound is the statement
    big, mid = R > 1e-3 * mx, R > 1e-6 * mx
    assert (D[big] <= tol * R[big]).all(), "per-entry relative error %.2e" % (D[big] / R[big]).max()
    assert (D[mid] <= 10 * tol * R[mid]).all(), "per-entry relative error %.2e" % (D[mid] / R[mid]).max()


def _blockwise_close(K, Kref, nvar, ndim, tol):
    n = K.shape[0]
    if nvar == ndim:
        _entrywise(K, Kref, tol)
        return
    mech = np.arange(n) % nvar != ndim
    for ra in (mech, ~mech):
        for ca in (mech, ~mech):
            _entrywise(K[ra][:, ca], Kref[ra][:, ca], tol)


def _vec_close(T, Tref, nvar, ndim, tol):
    if nvar == ndim:
        assert np.linalg.norm(T - Tref) <= tol * max(np.linalg.norm(Tref), 1e-300)
        return
    mech = np.arange(T.shape[0]) % nvar != ndim
    for m in (mech, ~mech):
        assert np.linalg.norm((T - Tref)[m]) <= tol * max(np.linalg.norm(Tref[m]), 1e-300)


@pytest.mark.parametrize("key", _cases())
def test_assembly_against_oracle_and_golden(key):
    from florence_b200 import backend
    from oracle import oracle as orc
    c = _load(key)
    matname = key.split("_", 3)[3]
    num = orc.MATERIAL_NUMBERS[matname]
    nnode, ndim = c["points"].shape
    electro = matname in ELEC
    nvar = ndim + (1 if electro else 0)
    form = 1 if electro else 0
    H = orc.hessian_size(num, ndim)
    n = nvar * nnode
    update = int(c["update"])
    h = backend.AssemblyHandle(c["points"], c["elements"], c["Jm"], c["AllGauss"], c["Bases"])
    mat = _material(backend, num, c["prm"])

    # --- sparsity pattern and slot maps: bit-exact against the reference's compiled ComputeSparsityPattern
    indices, indptr, dl, dg = h.sparsity_pattern(nvar, with_data_indices=True)
    assert indices.dtype == torch.int32 and indptr.dtype == torch.int32
    assert np.array_equal(indices.cpu().numpy(), c["sp_indices"])
    assert np.array_equal(indptr.cpu().numpy(), c["sp_indptr"])
    if c["sp_dl"] is not None:
        assert np.array_equal(dl.cpu().numpy(), c["sp_dl"])
        assert np.array_equal(dg.cpu().numpy(), c["sp_dg"])

    # --- implicit, COO mode
    Io, Jo, Vo, To = orc.assemble_implicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, H, update,
                                           c["prm"], num, mode="coo")
    I, J, V, T = h.assemble_implicit(c["Eulerx"], c["Eulerp"], mat, form, update, mode="coo")
    I, J, V, T = I.cpu().numpy(), J.cpu().numpy(), V.cpu().numpy(), T.cpu().numpy()
    assert np.array_equal(I, Io) and np.array_equal(J, Jo)
    Ko = csr_matrix((Vo, (Io, Jo)), shape=(n, n))
    K = csr_matrix((V, (I, J)), shape=(n, n))
    Kref = csr_matrix((c["K_data"], c["K_indices"], c["K_indptr"]), shape=(n, n))
    _blockwise_close(K, Ko, nvar, ndim, 1e-10)
    _blockwise_close(K, Kref, nvar, ndim, 1e-10)
    _vec_close(T, To, nvar, ndim, 1e-11)
    _vec_close(T, c["T"], nvar, ndim, 1e-11)
    # per-entry check of the raw element matrices against the oracle
    scale = np.abs(Vo).max() if not electro else None
    if not electro:
        assert np.abs(V - Vo).max() <= 1e-10 * scale

    # --- implicit, CSR mode (deterministic node-centric reduction) equals the oracle's slot-map mode
    pat = (c["sp_indices"], c["sp_indptr"], dl.cpu().numpy(), dg.cpu().numpy())
    V1o, T1o = orc.assemble_implicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, H, update,
                                     c["prm"], num, mode="csr", pattern=pat)
    V1, T1 = h.assemble_implicit(c["Eulerx"], c["Eulerp"], mat, form, update, mode="csr")
    V1b, _ = h.assemble_implicit(c["Eulerx"], c["Eulerp"], mat, form, update, mode="csr")
    assert torch.equal(V1, V1b), "CSR assembly must be bit-reproducible"
    K1 = csr_matrix((V1.cpu().numpy(), c["sp_indices"], c["sp_indptr"]), shape=(n, n))
    K1o = csr_matrix((V1o, c["sp_indices"], c["sp_indptr"]), shape=(n, n))
    _blockwise_close(K1, K1o, nvar, ndim, 1e-10)
    _blockwise_close(K1, Kref, nvar, ndim, 1e-10)
    _vec_close(T1.cpu().numpy(), T1o, nvar, ndim, 1e-11)
    # the CSR kernels (for hex64: dof-pair-plane scratch + wide reduction) against the scipy sum of the COO triplets of the same
    # state, which come from the element-major write path: only the summation order differs
    Kd = K.copy(); Kd.sum_duplicates(); Kd.sort_indices()
    _blockwise_close(K1, Kd, nvar, ndim, 1e-13)

    # --- explicit (matrix-free) internal force
    if update == 1:
        Teo = orc.assemble_explicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, c["prm"], num, form)
        Te = h.assemble_explicit(c["Eulerx"], c["Eulerp"], mat, form).cpu().numpy()
        _vec_close(Te, Teo, nvar, ndim, 1e-11)
        _vec_close(Te, c["T"], nvar, ndim, 1e-11)
    h.close()


def test_explicit_only_materials():
    """ExplicitMooneyRivlin (0) and ExplicitIsotropicElectroMechanics_108 (9): stress-only kernels of the explicit path."""
    from florence_b200 import backend
    from oracle import oracle as orc
    for key, num, form in (("asm_hex2_n2_MooneyRivlin", 0, 0), ("asm_quad1_n3_MooneyRivlin", 0, 0),
                           ("asm_hex2_n1_IsotropicElectroMechanics_108", 9, 1), ("asm_quad2_n2_IsotropicElectroMechanics_108", 9, 1)):
        c = _load(key)
        ndim = c["points"].shape[1]
        nvar = ndim + form
        h = backend.AssemblyHandle(c["points"], c["elements"], c["Jm"], c["AllGauss"], c["Bases"])
        Teo = orc.assemble_explicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, c["prm"], num, form)
        Te = h.assemble_explicit(c["Eulerx"], c["Eulerp"], _material(backend, num, c["prm"]), form).cpu().numpy()
        _vec_close(Te, Teo, nvar, ndim, 1e-11)
        h.close()


def test_unsupported_material_raises_not_implemented():
    from florence_b200 import backend
    c = _load("asm_hex1_n2_NeoHookean")
    h = backend.AssemblyHandle(c["points"], c["elements"], c["Jm"], c["AllGauss"], c["Bases"])
    with pytest.raises(NotImplementedError):
        h.assemble_explicit(c["Eulerx"], None, backend.make_material(7), 0)
    with pytest.raises(NotImplementedError):
        h.assemble_implicit(c["Eulerx"], None, backend.make_material(0), 0, True, mode="coo")
    with pytest.raises(ValueError):
        h.assemble_explicit(c["Eulerx"], None, backend.make_material(8, mu1=1., mu2=1., lamb=1., eps_2=1.), 0)
    h.close()


def test_laplacian_against_oracle_and_golden():
    from florence_b200 import backend
    from oracle import oracle as orc
    g = np.load(os.path.join(GOLD, "golden_laplacian.npz"))
    for key in [str(s) for s in g["lap_cases"]]:
        pts, els, Jm, AG, e = g[key + "_points"], g[key + "_elements"], g[key + "_Jm"], g[key + "_AllGauss"], g[key + "_e"]
        n = pts.shape[0]
        h = backend.AssemblyHandle(pts, els, Jm, AG)
        Io, Jo, Vo = orc.assemble_laplacian(pts, els, Jm, AG, -e, True, mode="coo")
        I, J, V = h.assemble_laplacian(-e, True, mode="coo")
        assert np.array_equal(I.cpu().numpy(), Io) and np.array_equal(J.cpu().numpy(), Jo)
        assert np.abs(V.cpu().numpy() - Vo).max() <= 1e-10 * np.abs(Vo).max(), key
        Kref = csr_matrix((g[key + "_K_data"], g[key + "_K_indices"], g[key + "_K_indptr"]), shape=(n, n))
        K = csr_matrix((V.cpu().numpy(), (Io, Jo)), shape=(n, n))
        assert abs(K - Kref).max() <= 1e-10 * abs(Kref).max(), key
        # non-symmetric branch and CSR mode
        en = -e + 0.1 * np.triu(np.ones_like(e), 1)
        Io, Jo, Vo = orc.assemble_laplacian(pts, els, Jm, AG, en, False, mode="coo")
        V2 = h.assemble_laplacian(en, False, mode="coo")[2].cpu().numpy()
        assert np.abs(V2 - Vo).max() <= 1e-10 * np.abs(Vo).max(), key
        indices, indptr = h.sparsity_pattern(1)
        Vc = h.assemble_laplacian(-e, True, mode="csr").cpu().numpy()
        Kc = csr_matrix((Vc, indices.cpu().numpy(), indptr.cpu().numpy()), shape=(n, n))
        assert abs(Kc - Kref).max() <= 1e-10 * abs(Kref).max(), key
        h.close()


def test_mass_lumped_and_consistent():
    from florence_b200 import backend
    from oracle import oracle as orc
    g = np.load(os.path.join(GOLD, "golden_explicit.npz"))
    pts, els, Jm, AG, Bases = g["exp_points"], g["exp_elements"], g["exp_Jm"], g["exp_AllGauss"], g["exp_Bases"]
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases)
    M = h.assemble_mass(1100.0, 3, "lumped").cpu().numpy()
    assert np.abs(M - g["exp_M_lumped"]).max() <= 1e-12 * np.abs(M).max()
    Io, Jo, Vo = orc.assemble_mass(pts, els, Bases, Jm, AG, 3, 1100.0, "consistent")
    I, J, V = h.assemble_mass(1100.0, 3, "consistent", mode="coo")
    assert np.array_equal(I.cpu().numpy(), Io) and np.array_equal(J.cpu().numpy(), Jo)
    assert np.abs(V.cpu().numpy() - Vo).max() <= 1e-12 * np.abs(Vo).max()
    h.close()


def test_explicit_time_loop_against_reference_trajectory():
    """Device-resident central-difference loop against the trajectory the reference's own integrator produced."""
    from florence_b200 import backend
    g = np.load(os.path.join(GOLD, "golden_explicit.npz"))
    pts, els, Jm, AG, Bases = g["exp_points"], g["exp_elements"], g["exp_Jm"], g["exp_AllGauss"], g["exp_Bases"]
    prm, rho, dt, nsteps = g["exp_prm"], float(g["exp_rho"]), float(g["exp_dt"]), int(g["exp_nsteps"])
    nnode = pts.shape[0]
    ref = g["exp_TotalDisp"]
    NF, AD = g["exp_neumann"], g["exp_applied_dirichlet"]
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases)
    mat = _material(backend, 1, prm)
    dev = h.device
    M = h.assemble_mass(rho, 3, "lumped")
    fixed = torch.zeros(nnode * 3, dtype=torch.uint8, device=dev)
    fixed[torch.as_tensor(g["exp_columns_out"].astype(np.int64), device=dev)] = 1
    X = torch.as_tensor(pts, device=dev).reshape(-1)
    Eulerx = X.clone()
    T = h.assemble_explicit(Eulerx.reshape(nnode, 3), None, mat, 0)
    # start-up (ExplicitStructuralDynamicIntegrator.py:57-81), zero initial displacement and velocity
    A0 = (torch.as_tensor(NF[:, 0], device=dev) - T) / M
    U0 = torch.zeros_like(T)
    U00 = (dt ** 2 / 2.) * A0
    U00[fixed.bool()] = 0.0
    incd = torch.zeros_like(T)
    for inc in range(2, nsteps):
        fext = torch.as_tensor(np.ascontiguousarray(NF[:, inc - 1]), device=dev)
        incd.zero_()
        incd[torch.as_tensor(g["exp_columns_out"].astype(np.int64), device=dev)] = torch.as_tensor(np.ascontiguousarray(AD[:, inc - 1]), device=dev)
        status = h.explicit_steps(mat, dt, 1, inc, M, fext, fixed, incd, U0, U00, Eulerx, T, fext_scale0=1.0, fext_scale_step=0.0)
        assert status == 0
        U = (Eulerx - X).reshape(nnode, 3).cpu().numpy()
        assert np.abs(U - ref[:, :, inc]).max() <= 1e-10 * np.abs(ref).max(), inc
    h.close()


def test_multi_step_fused_loop_equals_single_steps():
    from florence_b200 import backend
    c = _load("asm_hex2_n2_NeoHookean")
    nnode = c["points"].shape[0]
    h = backend.AssemblyHandle(c["points"], c["elements"], c["Jm"], c["AllGauss"], c["Bases"])
    mat = _material(backend, 1, c["prm"])
    dev = h.device
    M = h.assemble_mass(1100.0, 3, "lumped")
    fixed = torch.zeros(nnode * 3, dtype=torch.uint8, device=dev)
    fixed[:9] = 1
    fext = torch.zeros(nnode * 3, dtype=torch.float64, device=dev)
    fext[-3:] = 50.0
    dt = 2e-5

    def run(chunks):
        Eulerx = torch.as_tensor(c["Eulerx"], device=dev).reshape(-1).clone()
        T = h.assemble_explicit(Eulerx.reshape(nnode, 3), None, mat, 0)
        U0 = torch.zeros_like(T); U00 = torch.zeros_like(T)
        inc = 2
        for n in chunks:
            assert h.explicit_steps(mat, dt, n, inc, M, fext, fixed, None, U0, U00, Eulerx, T, fext_scale0=0.0, fext_scale_step=0.01) == 0
            inc += n
        return Eulerx, T, U0, U00
    a = run([1] * 12)
    b = run([12])
    cc = run([5, 7])
    for x, y, z in zip(a, b, cc):
        assert torch.equal(x, y) and torch.equal(x, z)
    h.close()


@pytest.mark.parametrize("p,n", [(2, 5), (1, 7)])
def test_explicit_tensor_core_kernel_equals_scalar_kernel_and_oracle(p, n):
    """hex8 / hex27 mechanics run the DMMA formulation; it must agree with the scalar kernel and the oracle, including
    batches that are only partly filled (nelem not a multiple of the batch size)."""
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    pts, els = flmesh.box_hex_mesh(n, n - 1, n + 1, p=p)
    Bases, Jm, AG = flmesh.tables("hex", p)
    x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.05, seed=5)
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases)
    for num, prm in ((1, dict(mu=4e5, lamb=2e6)), (2, dict(mu1=2e5, mu2=1e5, lamb=2e6)), (0, dict(mu1=2e5, mu2=1e5, lamb=2e6)),
                     (3, dict(mu1=2e5, mu2=3e4, mu3=2.3e6)), (10, dict(mu=4e5, lamb=2e6))):
        mat = backend.make_material(num, 0.0, **prm)
        To = orc.assemble_explicit(pts.numpy(), els.numpy(), x.numpy(), None, Jm, AG, 3, orc.params(**prm), num)
        h.set_option(0, 1)
        T1 = h.assemble_explicit(x, None, mat, 0).cpu().numpy()
        h.set_option(0, 0)
        T0 = h.assemble_explicit(x, None, mat, 0).cpu().numpy()
        assert np.linalg.norm(T1 - To) <= 1e-11 * np.linalg.norm(To)
        assert np.linalg.norm(T0 - To) <= 1e-11 * np.linalg.norm(To)
        assert np.abs(T1 - T0).max() <= 1e-11 * np.abs(To).max()
    h.close()


@pytest.mark.parametrize("key", ["asm_hex2_n2_NeoHookean", "asm_hex3_n1_MooneyRivlin", "asm_hex2_n1_IsotropicElectroMechanics_108",
                                 "asm_hex3_n1_IsotropicElectroMechanics_108", "asm_hex2_n1_NearlyIncompressibleMooneyRivlin",
                                 "asm_tet3_n1_NeoHookean", "asm_tet3_n2x1x1_MooneyRivlin", "asm_tet3_n1_IsotropicElectroMechanics_108",
                                 "asm_tet3_n1_LinearElastic"])
def test_implicit_tensor_core_kernel_equals_generic_kernel(key):
    """hex64 runs the DMMA parent-space formulation by default, hex27 and tet20 (20 -> 32 node padding) when option 1 is 2; the
    generic column-owner kernel must give the same K_e, and both must match the reference's K."""
    from florence_b200 import backend
    from oracle import oracle as orc
    c = _load(key)
    matname = key.split("_", 3)[3]
    num = orc.MATERIAL_NUMBERS[matname]
    form = 1 if matname in ELEC else 0
    ndim = 3
    nvar = ndim + form
    h = backend.AssemblyHandle(c["points"], c["elements"], c["Jm"], c["AllGauss"], c["Bases"])
    mat = _material(backend, num, c["prm"])
    out = {}
    for opt, val in ((1, 2), (0, 0)):   # 2 = force the tensor-core kernel also for hex27
        h.set_option(1, val)
        I, J, V, T = h.assemble_implicit(c["Eulerx"], c["Eulerp"], mat, form, int(c["update"]), mode="coo")
        out[opt] = (V.cpu().numpy(), T.cpu().numpy(), I.cpu().numpy(), J.cpu().numpy())
    n = nvar * c["points"].shape[0]
    K1 = csr_matrix((out[1][0], (out[1][2], out[1][3])), shape=(n, n))
    K0 = csr_matrix((out[0][0], (out[0][2], out[0][3])), shape=(n, n))
    _blockwise_close(K1, K0, nvar, ndim, 1e-10)
    _vec_close(out[1][1], out[0][1], nvar, ndim, 1e-11)
    Kref = csr_matrix((c["K_data"], c["K_indices"], c["K_indptr"]), shape=(n, n))
    _blockwise_close(K1, Kref, nvar, ndim, 1e-10)
    _vec_close(out[1][1], c["T"], nvar, ndim, 1e-11)
    # CSR mode through the tensor-core kernel: the same matrix, deterministic
    h.set_option(1, 2)
    V1, _ = h.assemble_implicit(c["Eulerx"], c["Eulerp"], mat, form, int(c["update"]), mode="csr")
    V1b, _ = h.assemble_implicit(c["Eulerx"], c["Eulerp"], mat, form, int(c["update"]), mode="csr")
    assert torch.equal(V1, V1b)
    Kc = csr_matrix((V1.cpu().numpy(), c["sp_indices"], c["sp_indptr"]), shape=(n, n))
    _blockwise_close(Kc, Kref, nvar, ndim, 1e-10)
    h.close()


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 2, 2)])
@pytest.mark.parametrize("matname", ["IsotropicElectroMechanics_108", "MooneyRivlin"])
def test_hex64_multi_element_against_oracle(matname, shape):
    """Config 4's kernels on meshes where interior nodes are visited by up to 8 elements: the DMMA element
    kernel with its dof-pair-plane scratch and csr_gather_wide_kernel<NV,64> against the oracle, CSR and COO, and CSR against the
    scipy sum of the COO triplets."""
    from florence_b200 import backend, mesh as flmesh
    from oracle import oracle as orc
    nx, ny, nz = shape
    pts, els = flmesh.box_hex_mesh(nx, ny, nz, p=3, lengths=(1.0, 0.8, 1.2))
    Bases, Jm, AG = flmesh.tables("hex", 3)
    x = flmesh.perturbed_state(pts, 1.0 / (3 * nx), 0.03, seed=17)
    electro = matname in ELEC
    num = orc.MATERIAL_NUMBERS[matname]
    nvar, form, H = (4, 1, 9) if electro else (3, 0, 6)
    rng = np.random.default_rng(3)
    phi = (9.0e3 * pts[:, 2].numpy() + 10.0 * rng.uniform(-1, 1, pts.shape[0])) if electro else None
    prm = dict(mu1=2.4e5, mu2=1.6e5, lamb=2.0e6)
    if electro:
        prm["eps_2"] = 4.0 * 8.8541e-12
    P, E, X = pts.numpy(), els.numpy(), x.numpy()
    n = nvar * P.shape[0]
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases)
    mat = backend.make_material(num, 0.0, **prm)
    pat = orc.sparsity_pattern(E, P.shape[0], nvar)
    indices, indptr = h.sparsity_pattern(nvar)
    assert np.array_equal(indices.cpu().numpy(), pat[0]) and np.array_equal(indptr.cpu().numpy(), pat[1])
    Io, Jo, Vo, To = orc.assemble_implicit(P, E, X, phi, Jm, AG, nvar, H, 1, orc.params(**prm), num, mode="coo")
    Ko = csr_matrix((Vo, (Io, Jo)), shape=(n, n)); Ko.sum_duplicates(); Ko.sort_indices()
    for opt in (1, 0):          # default (DMMA for hex64) and the generic column-owner kernel
        h.set_option(1, opt)
        I, J, V, T = h.assemble_implicit(x, phi, mat, form, True, mode="coo")
        assert np.array_equal(I.cpu().numpy(), Io) and np.array_equal(J.cpu().numpy(), Jo)
        K = csr_matrix((V.cpu().numpy(), (Io, Jo)), shape=(n, n)); K.sum_duplicates(); K.sort_indices()
        _blockwise_close(K, Ko, nvar, 3, 1e-10)
        _vec_close(T.cpu().numpy(), To, nvar, 3, 1e-11)
        V1, T1 = h.assemble_implicit(x, phi, mat, form, True, mode="csr")
        V1b, _ = h.assemble_implicit(x, phi, mat, form, True, mode="csr")
        assert torch.equal(V1, V1b)
        K1 = csr_matrix((V1.cpu().numpy(), pat[0], pat[1]), shape=(n, n))
        _blockwise_close(K1, Ko, nvar, 3, 1e-10)
        _blockwise_close(K1, K, nvar, 3, 1e-13)
        _vec_close(T1.cpu().numpy(), To, nvar, 3, 1e-11)
    h.close()


@pytest.mark.parametrize("electro", [True, False])
def test_hex64_pipelined_reduction_is_bit_identical(electro):
    """The cross-node software-pipelined CSR reduction of hex64 (csr_gather_hex64_kernel: nvar 4 electro-mechanics, nvar 3
    mechanics) adds the same visits in the same order as the block-per-node kernel it replaces (option 5 = 1): V must not differ
    in a single bit, on a mesh with every node class (vertex nodes with 8 visits down to interior nodes with one) and for both
    plane-major scratch layouts."""
    from florence_b200 import backend, mesh as flmesh
    pts, els = flmesh.box_hex_mesh(3, 3, 2, p=3)
    Bases, Jm, AG = flmesh.tables("hex", 3)
    x = flmesh.perturbed_state(pts, 1.0 / 9, 0.02, seed=4)
    gen = torch.Generator(); gen.manual_seed(9)
    phi = 9e3 * pts[:, 2] + 10.0 * (2 * torch.rand(pts.shape[0], dtype=torch.float64, generator=gen) - 1)
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases)
    mu = 5e4
    mat = backend.make_material(8 if electro else 2, 1200.0, mu1=mu, mu2=mu, lamb=2 * mu * 0.4 / (1 - 0.8), eps_2=4 * 8.8541e-12)
    if not electro:
        phi = None
    out = {}
    for layout in (1, 3):                      # per-row-node planes (default) / dof-pair planes
        h.set_option(1, layout)
        for unpipelined in (0, 1):
            h.set_option(5, unpipelined)
            V, T = h.assemble_implicit(x, phi, mat, 1 if electro else 0, True, mode="csr")
            out[(layout, unpipelined)] = V.clone()
        assert torch.equal(out[(layout, 0)], out[(layout, 1)])
    assert torch.equal(out[(1, 0)], out[(3, 0)])
    h.close()
