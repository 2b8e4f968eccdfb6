"""Test double for backend.AssemblyHandle on hosts without a GPU: same methods and return shapes, numbers from the CPU oracle.

Only tests import this (the product path has no CPU fallback).  It lets the HOST logic of the plug-in layer -- name lookup,
argument unpacking from Florence objects, COO/CSR wrapping, the parallel launchers -- run in `-m "not gpu"`."""
import numpy as np
import torch

from oracle import oracle as orc


class FakeHandle(object):
    device = torch.device("cpu")

    def __init__(self, points, elements, Jm, AllGauss, Bases=None, device=None):
        self.P = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
        self.E = np.ascontiguousarray(np.asarray(elements).astype(np.int64))
        self.Jm, self.AG, self.Bases = np.asarray(Jm), np.asarray(AllGauss), Bases
        self.nnode, self.ndim = self.P.shape
        self.nelem, self.npe = self.E.shape
        self._pat = {}
        self.calls = []

    def close(self):
        pass

    def _prm(self, m):
        return orc.params(mu=m.mu, mu1=m.mu1, mu2=m.mu2, mu3=m.mu3, mue=m.mue, lamb=m.lamb, eps_1=m.eps_1, eps_2=m.eps_2, eps_3=m.eps_3,
                          eps_e=m.eps_e)

    def build_pattern(self, nvar):
        if nvar not in self._pat:
            self._pat[nvar] = orc.sparsity_pattern(self.E, self.nnode, nvar)
        return self._pat[nvar][0].shape[0]

    def sparsity_pattern(self, nvar, with_data_indices=False):
        self.build_pattern(nvar)
        pat = self._pat[nvar]
        out = pat if with_data_indices else pat[:2]
        return tuple(torch.as_tensor(a) for a in out)

    def assemble_implicit(self, Eulerx, Eulerp, material, formulation_number=0, requires_geometry_update=True, mode="csr", out=None,
                          with_indices=True):
        self.calls.append(("implicit", mode))
        nvar = self.ndim + (1 if formulation_number == 1 else 0)
        num = material.material_number
        H = orc.hessian_size(num, self.ndim)
        x = np.asarray(Eulerx, dtype=np.float64).reshape(self.nnode, self.ndim)
        p = None if Eulerp is None else np.asarray(Eulerp, dtype=np.float64).reshape(-1)
        if mode == "coo":
            I, J, V, T = orc.assemble_implicit(self.P, self.E, x, p, self.Jm, self.AG, nvar, H, int(requires_geometry_update),
                                               self._prm(material), num, mode="coo")
            return tuple(torch.as_tensor(a) for a in (I, J, V, T))
        self.build_pattern(nvar)
        V, T = orc.assemble_implicit(self.P, self.E, x, p, self.Jm, self.AG, nvar, H, int(requires_geometry_update), self._prm(material), num,
                                     mode="csr", pattern=self._pat[nvar])
        return torch.as_tensor(V), torch.as_tensor(T)

    def assemble_explicit(self, Eulerx, Eulerp, material, formulation_number=0, out=None):
        self.calls.append(("explicit",))
        nvar = self.ndim + (1 if formulation_number == 1 else 0)
        x = np.asarray(Eulerx, dtype=np.float64).reshape(self.nnode, self.ndim)
        p = None if Eulerp is None else np.asarray(Eulerp, dtype=np.float64).reshape(-1)
        T = orc.assemble_explicit(self.P, self.E, x, p, self.Jm, self.AG, nvar, self._prm(material), material.material_number, formulation_number)
        return torch.as_tensor(T)

    def assemble_laplacian(self, e_tensor, is_hessian_symmetric=True, mode="csr"):
        self.calls.append(("laplacian", mode))
        I, J, V = orc.assemble_laplacian(self.P, self.E, self.Jm, self.AG, np.asarray(e_tensor, dtype=np.float64), bool(is_hessian_symmetric),
                                         mode="coo")
        if mode == "coo":
            return tuple(torch.as_tensor(a) for a in (I, J, V))
        from scipy.sparse import csr_matrix
        K = csr_matrix((V, (I, J)), shape=(self.nnode, self.nnode))
        K.sum_duplicates(); K.sort_indices()
        return torch.as_tensor(K.data)

    def assemble_mass(self, rho, nvar, mass_type="lumped", mode="coo"):
        self.calls.append(("mass", mass_type))
        out = orc.assemble_mass(self.P, self.E, self.Bases, self.Jm, self.AG, nvar, float(rho), mass_type)
        if mass_type == "lumped":
            return torch.as_tensor(out)
        return tuple(torch.as_tensor(a) for a in out)

    def row_block(self, nvar, V, owned_nodes, node_map):
        """Host restatement of fl_row_block_build / fl_row_block_emit."""
        idx, iptr = self._pat[nvar][:2]
        V = np.asarray(V)
        own, gl = np.asarray(owned_nodes), np.asarray(node_map)
        rows = (own[:, None] * nvar + np.arange(nvar)[None, :]).ravel()
        cnt = iptr[rows + 1] - iptr[rows]
        take = np.concatenate([np.arange(iptr[r], iptr[r + 1]) for r in rows]) if rows.size else np.zeros(0, np.int64)
        cl = idx[take]
        cols = gl[cl // nvar] * nvar + cl % nvar
        return torch.as_tensor(np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)), torch.as_tensor(cols.astype(np.int64)), torch.as_tensor(V[take])


def install_fake(monkeypatch):
    """Route florence_b200.assembly through FakeHandle (CPU tensors, plain copies instead of pinned D2H)."""
    from florence_b200 import assembly
    monkeypatch.setattr(assembly, "AssemblyHandle", FakeHandle)
    monkeypatch.setattr(assembly, "_to_host", lambda t, tag, defer=False, prepared=None: t.numpy().copy())
    monkeypatch.setattr(assembly, "_to_host_many", lambda items, prepared=None: tuple(t.numpy().copy() for t, _ in items))
    assembly._handle_cache.clear()
    return assembly
