"""CPU restatement (numpy + scipy, the libraries the reference itself uses here) of the Dirichlet reduction.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu legs, never by the product.

Follows Florence/BoundaryCondition/BoundaryCondition.py:
  columns_in                          :396   np.delete(arange(nvar*nnode), columns_out)
  GetReducedMatrices                  :842-858
  ApplyDirichletGetReducedMatrices    :861-891
  UpdateFixDoFs / UpdateFreeDoFs      :908-932
Pinned against the reference's own methods by tests/golden/make_golden_dirichlet.py -> tests/golden/golden_dirichlet.npz.
"""
import numpy as np
from scipy.sparse import csr_matrix


def columns_in(ndof_total, columns_out):
    """:396"""
    return np.delete(np.arange(0, ndof_total), columns_out)


def get_reduced_matrices(stiffness, F, cols_in, only_residual=False):
    """:842-858.  F has shape (N, 1)."""
    F_b = F[cols_in, 0]
    if only_residual:
        return F_b
    stiffness_b = stiffness[cols_in, :][:, cols_in]
    return stiffness_b, F_b


def apply_dirichlet_get_reduced_matrices(stiffness, F, applied, cols_in, cols_out, load_factor=1.0, mass=None, only_residual=False):
    """:861-891.  F (N, 1) is modified in place, exactly like the reference."""
    nnz_cols = ~np.isclose(applied, 0.0)
    F[cols_in] = F[cols_in] - (stiffness[cols_in, :][:, cols_out[nnz_cols]] * applied[nnz_cols] * load_factor)[:, None]
    if only_residual:
        return F
    F_b = F[cols_in, 0]
    stiffness_b = stiffness[cols_in, :][:, cols_in]
    if mass is not None:
        mass_b = mass[cols_in, :][:, cols_in]
        return stiffness_b, F_b, F, mass_b
    return stiffness_b, F_b, F


def update_fix_dofs(applied_inc, cols_out, fsize, nvar):
    """:908-919"""
    total = np.zeros((fsize, 1))
    total[cols_out, 0] = applied_inc
    return total.reshape(fsize // nvar, nvar)


def update_free_dofs(sol, cols_in, fsize, nvar):
    """:921-932"""
    total = np.zeros((fsize, 1))
    total[cols_in, 0] = sol
    return total.reshape(fsize // nvar, nvar)


def full_csr(V, indices, indptr, n):
    return csr_matrix((np.asarray(V), np.asarray(indices), np.asarray(indptr)), shape=(n, n), dtype=np.float64)
