"""Device-resident mirror of the Dirichlet reduction of Florence/BoundaryCondition/BoundaryCondition.py.

The reference slices `K[columns_in,:][:,columns_in]` with scipy on the host in every Newton iteration
(GetReducedMatrices :842-858, ApplyDirichletGetReducedMatrices :861-891; callers Florence/Solver/FEMSolver.py:769, :951).  Here the
K values stay on the GPU in the CSR order of ComputeSparsityPattern, the reduced pattern is built once per boundary condition
(fl_dirichlet_build) and every iteration is one kernel pass over the free rows (fl_dirichlet_apply) -- the 2.8 GB D2H of K per
assembly of the tet10 benchmark disappears when a device solver consumes K_b.

Method names, argument meaning and return order follow the reference; arrays are torch device tensors:
  stiffness / mass : CSR values aligned with AssemblyHandle.sparsity_pattern(nvar) (what _LowLevelAssembly_ produces)
  F                : (N,) or (N,1) float64, modified in place exactly where the reference modifies it
  returns          : ReducedCSR(data, indices, indptr, shape) for matrices, 1-D tensors for vectors
"""
import collections

import numpy as np
import torch

from . import backend

ReducedCSR = collections.namedtuple("ReducedCSR", ["data", "indices", "indptr", "shape"])


def _to_scipy(r):
    from scipy.sparse import csr_matrix
    return csr_matrix((r.data.cpu().numpy(), r.indices.cpu().numpy(), r.indptr.cpu().numpy()), shape=r.shape)


ReducedCSR.to_scipy = _to_scipy


class DeviceBoundaryCondition(object):
    """columns_out / columns_in / applied_dirichlet as BoundaryCondition.GetDirichletBoundaryConditions leaves them (:375-396)."""

    def __init__(self, handle, nvar, columns_out, analysis_type="static"):
        self.handle, self.nvar = handle, nvar
        co = np.asarray(columns_out.cpu() if isinstance(columns_out, torch.Tensor) else columns_out).astype(np.int64).ravel()
        self.fsize = handle.nnode * nvar
        self.n_in, self.nnz_b = handle.dirichlet_build(nvar, co)
        self.indices_b, self.indptr_b, cin = handle.dirichlet_pattern()
        self.columns_out = torch.as_tensor(co, device=handle.device)
        self.columns_in = cin.long()
        self.analysis_type = analysis_type

    @classmethod
    def from_flags(cls, handle, dirichlet_flags, analysis_type="static"):
        """dirichlet_flags (nnode x nvar), NaN = free: columns_out and applied_dirichlet as BoundaryCondition.py:391-395."""
        flat = np.asarray(dirichlet_flags, dtype=np.float64).ravel()
        mask = ~np.isnan(flat)
        self = cls(handle, np.asarray(dirichlet_flags).shape[1], np.arange(flat.size)[mask], analysis_type)
        self.applied_dirichlet = torch.as_tensor(flat[mask], device=handle.device)
        return self

    def _csr(self, data):
        return ReducedCSR(data, self.indices_b, self.indptr_b, (self.n_in, self.n_in))

    @staticmethod
    def _flat(F):
        f = F.reshape(-1)
        if f.data_ptr() != F.data_ptr():
            raise ValueError("F must be contiguous")
        return f

    def GetReducedMatrices(self, stiffness, F, mass=None, only_residual=False):
        """BoundaryCondition.py:842-858 -> (stiffness_b, F_b, mass_b)."""
        f = self._flat(F)
        if only_residual:
            return self.handle.dirichlet_apply(None, f, None, want_values=False)[1]
        Vb, Fb = self.handle.dirichlet_apply(stiffness, f, None)
        return self._csr(Vb), Fb, torch.empty(0, dtype=torch.float64, device=f.device)

    def ApplyDirichletGetReducedMatrices(self, stiffness, F, AppliedDirichlet, LoadFactor=1., mass=None, only_residual=False):
        """BoundaryCondition.py:861-891 -> (stiffness_b, F_b, F[, mass_b]); F is updated in place."""
        f = self._flat(F)
        if only_residual:
            self.handle.dirichlet_apply(stiffness, f, AppliedDirichlet, LoadFactor, want_values=False, want_reduced_force=False)
            return F
        Vb, Fb = self.handle.dirichlet_apply(stiffness, f, AppliedDirichlet, LoadFactor)
        if self.analysis_type != 'static':
            Mb, _ = self.handle.dirichlet_apply(mass, None, None, want_reduced_force=False)
            return self._csr(Vb), Fb, F, self._csr(Mb)
        return self._csr(Vb), Fb, F

    def UpdateFixDoFs(self, AppliedDirichletInc, fsize=None, nvar=None):
        """:908-919"""
        nvar = nvar or self.nvar
        total = torch.zeros(fsize or self.fsize, dtype=torch.float64, device=self.columns_out.device)
        total[self.columns_out] = backend.to_device(AppliedDirichletInc, torch.float64, total.device).reshape(-1)
        return total.reshape(-1, nvar)

    def UpdateFreeDoFs(self, sol, fsize=None, nvar=None):
        """:921-932"""
        nvar = nvar or self.nvar
        total = torch.zeros(fsize or self.fsize, dtype=torch.float64, device=self.columns_in.device)
        total[self.columns_in] = backend.to_device(sol, torch.float64, total.device).reshape(-1)
        return total.reshape(-1, nvar)
