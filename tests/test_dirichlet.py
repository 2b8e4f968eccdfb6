"""Dirichlet reduction: oracle vs the reference's golden vectors on CPU, device vs both on the GPU."""
import os

import numpy as np
import pytest

from oracle import dirichlet as od

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_dirichlet.npz"))
CASES = ["hex8", "tet10"]


def _case(tag):
    from oracle import oracle
    g = {k[len(tag) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(tag + "_")}
    nvar = int(g["nvar"])
    indices, indptr = oracle.sparsity_pattern(g["elements"], g["points"].shape[0], nvar, with_data_indices=False)
    return g, nvar, indices, indptr


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_golden(tag):
    g, nvar, indices, indptr = _case(tag)
    N = nvar * g["points"].shape[0]
    K = od.full_csr(g["V"], indices, indptr, N)
    cols_out = g["columns_out"]
    cols_in = od.columns_in(N, cols_out)
    F = g["F"].copy()[:, None]
    Kb, Fb = od.get_reduced_matrices(K, F, cols_in)
    assert np.array_equal(Kb.data, g["Kb_data"]) and np.array_equal(Kb.indices, g["Kb_indices"]) and np.array_equal(Kb.indptr, g["Kb_indptr"])
    assert np.array_equal(Fb, g["Fb_plain"])
    Kb2, Fb2, Fm = od.apply_dirichlet_get_reduced_matrices(K, F, g["applied"], cols_in, cols_out, float(g["load_factor"]))
    assert np.array_equal(Fb2, g["Fb_applied"]) and np.array_equal(Fm[:, 0], g["F_applied"]) and np.array_equal(Kb2.data, g["Kb2_data"])
    assert np.array_equal(od.update_fix_dofs(g["applied"], cols_out, N, nvar), g["dU_fix"])
    assert np.array_equal(od.update_free_dofs(Fb2, cols_in, N, nvar), g["dU_free"])


# ------------------------------------------------------------------------------------------------------------------ GPU
def _device_case(tag):
    import torch
    from florence_b200 import backend, boundary, mesh as flmesh
    g, nvar, indices, indptr = _case(tag)
    kind, p = ("hex", 1) if tag == "hex8" else ("tet", 2)
    B, Jm, AG = flmesh.tables(kind, p)
    dev = torch.device("cuda:0")
    h = backend.AssemblyHandle(g["points"], g["elements"].astype(np.uint64), Jm, AG, B, device=dev)
    h.build_pattern(nvar)
    bc = boundary.DeviceBoundaryCondition(h, nvar, g["columns_out"])
    return g, nvar, h, bc, dev


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_device_reduction_bit_exact(tag):
    import torch
    g, nvar, h, bc, dev = _device_case(tag)
    N = nvar * g["points"].shape[0]
    assert bc.n_in == N - g["columns_out"].shape[0] and bc.nnz_b == g["Kb_data"].shape[0]
    assert np.array_equal(bc.indptr_b.cpu().numpy(), g["Kb_indptr"])
    assert np.array_equal(bc.indices_b.cpu().numpy(), g["Kb_indices"])
    assert np.array_equal(bc.columns_in.cpu().numpy(), od.columns_in(N, g["columns_out"]))
    V = torch.as_tensor(g["V"], device=dev)
    F = torch.as_tensor(g["F"].copy(), device=dev).reshape(-1, 1)
    Kb, Fb, Mb = bc.GetReducedMatrices(V, F)
    assert np.array_equal(Kb.data.cpu().numpy(), g["Kb_data"]) and np.array_equal(Fb.cpu().numpy(), g["Fb_plain"])
    assert np.array_equal(F.cpu().numpy()[:, 0], g["F"])                                    # untouched
    assert np.array_equal(bc.GetReducedMatrices(V, F, only_residual=True).cpu().numpy(), g["Fb_plain"])
    Kb2, Fb2, Fm = bc.ApplyDirichletGetReducedMatrices(V, F, g["applied"], LoadFactor=float(g["load_factor"]))
    assert Fm is F
    assert np.array_equal(Kb2.data.cpu().numpy(), g["Kb2_data"])
    assert np.array_equal(Fb2.cpu().numpy(), g["Fb_applied"])
    assert np.array_equal(F.cpu().numpy()[:, 0], g["F_applied"])
    assert np.array_equal(bc.UpdateFixDoFs(g["applied"]).cpu().numpy(), g["dU_fix"])
    assert np.array_equal(bc.UpdateFreeDoFs(Fb2).cpu().numpy(), g["dU_free"])
    # only_residual returns the full, modified F
    F2 = torch.as_tensor(g["F"].copy(), device=dev).reshape(-1, 1)
    out = bc.ApplyDirichletGetReducedMatrices(V, F2, g["applied"], LoadFactor=float(g["load_factor"]), only_residual=True)
    assert np.array_equal(out.cpu().numpy()[:, 0], g["F_applied"])
    # dynamic analysis also reduces the mass with the same pattern
    bc.analysis_type = "dynamic"
    F3 = torch.as_tensor(g["F"].copy(), device=dev)
    r = bc.ApplyDirichletGetReducedMatrices(V, F3, g["applied"], LoadFactor=float(g["load_factor"]), mass=2.0 * V)
    assert len(r) == 4 and np.array_equal(r[3].data.cpu().numpy(), 2.0 * g["Kb_data"])
    h.close()


@pytest.mark.gpu
def test_device_reduction_edge_cases():
    import torch
    from florence_b200 import backend, boundary, mesh as flmesh
    dev = torch.device("cuda:0")
    pts, els = flmesh.box_hex_mesh(2, 2, 2, p=1)
    B, Jm, AG = flmesh.tables("hex", 1)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    nnz = h.build_pattern(3)
    N = 3 * pts.shape[0]
    V = torch.arange(nnz, dtype=torch.float64, device=dev)
    F = torch.arange(N, dtype=torch.float64, device=dev)
    # no prescribed dof: identity
    bc = boundary.DeviceBoundaryCondition(h, 3, np.zeros(0, np.int64))
    Kb, Fb, _ = bc.GetReducedMatrices(V, F)
    assert bc.n_in == N and torch.equal(Kb.data, V) and torch.equal(Fb, F)
    # every dof prescribed: empty system
    bc = boundary.DeviceBoundaryCondition(h, 3, np.arange(N))
    Kb, Fb, _ = bc.GetReducedMatrices(V, F)
    assert bc.n_in == 0 and Kb.data.numel() == 0 and Fb.numel() == 0 and Kb.indptr.cpu().tolist() == [0]
    # unsorted / out-of-range lists are rejected like an index error, not silently reordered
    with pytest.raises(ValueError):
        boundary.DeviceBoundaryCondition(h, 3, np.array([5, 2]))
    with pytest.raises(ValueError):
        boundary.DeviceBoundaryCondition(h, 3, np.array([N]))
    # apply before build on a fresh handle is a state error
    h2 = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    h2.build_pattern(3)
    h2._dirichlet = (3, N, nnz, 0)
    with pytest.raises(Exception):
        h2.dirichlet_apply(V, F, None)
    h.close(); h2.close()


@pytest.mark.gpu
def test_device_reduction_full_size_properties():
    """Config-2-sized system (n=24 here keeps the scipy check in seconds): reduced K equals scipy's slicing of the device K."""
    import torch
    from florence_b200 import backend, boundary, mesh as flmesh
    dev = torch.device("cuda:0")
    n = 12
    pts, els = flmesh.box_tet_mesh(n, n, n, p=2, device=dev)
    B, Jm, AG = flmesh.tables("tet", 2)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    h.build_pattern(3)
    x = flmesh.perturbed_state(pts, 1.0 / n, 1e-3 * n, seed=3)
    V, T = h.assemble_implicit(x, None, backend.make_material(10, 0.0, mu=1e5, lamb=1.5e5), 0, True, mode="csr")
    flags = np.full((pts.shape[0], 3), np.nan)
    z = pts[:, 2].cpu().numpy()
    flags[np.isclose(z, 0.0)] = 0.0
    flags[np.isclose(z, z.max()), 2] = 0.01
    bc = boundary.DeviceBoundaryCondition.from_flags(h, flags)
    F = T.clone().reshape(-1, 1)
    Kb, Fb, Fm = bc.ApplyDirichletGetReducedMatrices(V, F, bc.applied_dirichlet, LoadFactor=0.5)
    indices, indptr = h.sparsity_pattern(3)
    N = 3 * pts.shape[0]
    K = od.full_csr(V.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy(), N)
    Fh = T.cpu().numpy().copy()[:, None]
    cols_out = bc.columns_out.cpu().numpy()
    Kb_o, Fb_o, F_o = od.apply_dirichlet_get_reduced_matrices(K, Fh, bc.applied_dirichlet.cpu().numpy(), od.columns_in(N, cols_out), cols_out, 0.5)
    assert np.array_equal(Kb.data.cpu().numpy(), Kb_o.data) and np.array_equal(Kb.indices.cpu().numpy(), Kb_o.indices)
    assert np.array_equal(Fb.cpu().numpy(), Fb_o) and np.array_equal(F.cpu().numpy(), F_o)
    # the reduced matrix of a symmetric K is symmetric
    S = Kb.to_scipy()
    assert abs(S - S.T).max() <= 1e-9 * abs(S).max()
    h.close()
