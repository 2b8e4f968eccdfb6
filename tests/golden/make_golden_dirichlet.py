"""Golden vectors of the Dirichlet reduction, produced by the REFERENCE's own BoundaryCondition methods
(Florence/BoundaryCondition/BoundaryCondition.py:842-891, :908-932) on a small assembled system.

Run in the build container only:  python tests/golden/make_golden_dirichlet.py
The stiffness fed to the reference is the oracle's CSR assembly of a 3x3x3 hex8 NeoHookean mesh (any sparse matrix with the
mesh's pattern would do: the reduction is independent of how K was produced).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _load_reference  # noqa: E402

Florence = _load_reference.load()
from Florence.BoundaryCondition.BoundaryCondition import BoundaryCondition  # noqa: E402

from florence_b200 import mesh as flmesh  # noqa: E402
from oracle import oracle  # noqa: E402
from oracle import dirichlet as od  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    out = {}
    for tag, kind, p, n, nvar, matnum in (("hex8", "hex", 1, 3, 3, 1), ("tet10", "tet", 2, 2, 3, 10)):
        if kind == "hex":
            pts, els = flmesh.box_hex_mesh(n, n, n, p=p)
        else:
            pts, els = flmesh.box_tet_mesh(n, n, n, p=p)
        pts, els = pts.numpy(), els.numpy()
        B, Jm, AG = flmesh.tables(kind, p)
        x = pts + 0.02 / (p * n) * rng.uniform(-1, 1, pts.shape)
        pattern = oracle.sparsity_pattern(els, pts.shape[0], nvar)
        indices, indptr = pattern[0], pattern[1]
        prm = oracle.params(mu=1e5, lamb=1.5e5)
        V, T = oracle.assemble_implicit(pts, els, x, None, Jm, AG, nvar, 6, 1, prm, matnum, mode="csr", pattern=pattern)
        N = nvar * pts.shape[0]
        K = od.full_csr(V, indices, indptr, N)
        # bottom face clamped, top face moved in z by a non-zero amount, a few exact zeros and tiny values among the applied dofs
        zmin, zmax = pts[:, 2].min(), pts[:, 2].max()
        flags = np.full((pts.shape[0], nvar), np.nan)
        flags[np.isclose(pts[:, 2], zmin)] = 0.0
        top = np.isclose(pts[:, 2], zmax)
        flags[top, 2] = 0.05
        flags[top, 0] = 1e-9            # below np.isclose's atol: dropped by the reference
        flat = flags.ravel()
        cols_out = np.arange(flat.size)[~np.isnan(flat)].astype(np.int64)      # BoundaryCondition.py:391-395
        cols_in = np.delete(np.arange(0, N), cols_out)                         # :396
        applied = flat[~np.isnan(flat)]
        bc = BoundaryCondition()
        bc.columns_in, bc.columns_out = cols_in, cols_out
        bc.analysis_type = "static"
        F = rng.standard_normal((N, 1))
        Kb, Fb, _ = bc.GetReducedMatrices(K, F.copy())
        F2 = F.copy()
        Kb2, Fb2, Fmod = bc.ApplyDirichletGetReducedMatrices(K, F2, applied, LoadFactor=0.35)
        assert Kb.has_canonical_format or True
        dU_fix = bc.UpdateFixDoFs(applied, N, nvar)
        dU_free = bc.UpdateFreeDoFs(Fb2, N, nvar)
        out.update({tag + "_points": pts, tag + "_elements": els.astype(np.int64), tag + "_V": V, tag + "_F": F[:, 0],
                    tag + "_columns_out": cols_out, tag + "_applied": applied, tag + "_load_factor": np.array(0.35),
                    tag + "_Kb_data": Kb.data, tag + "_Kb_indices": Kb.indices, tag + "_Kb_indptr": Kb.indptr,
                    tag + "_Fb_plain": Fb, tag + "_Fb_applied": Fb2, tag + "_F_applied": Fmod[:, 0],
                    tag + "_Kb2_data": Kb2.data, tag + "_dU_fix": dU_fix, tag + "_dU_free": dU_free,
                    tag + "_nvar": np.array(nvar), tag + "_p": np.array(p), tag + "_n": np.array(n)})
        # the restatement against the reference, bit for bit
        Kb_o, Fb_o = od.get_reduced_matrices(K, F.copy(), cols_in)
        assert np.array_equal(Kb_o.data, Kb.data) and np.array_equal(Fb_o, Fb)
        F3 = F.copy()
        Kb3, Fb3, F3m = od.apply_dirichlet_get_reduced_matrices(K, F3, applied, cols_in, cols_out, 0.35)
        assert np.array_equal(Fb3, Fb2) and np.array_equal(F3m, Fmod) and np.array_equal(Kb3.indices, Kb2.indices)
    np.savez_compressed(os.path.join(HERE, "golden_dirichlet.npz"), **out)
    print("wrote golden_dirichlet.npz", {k: v.shape for k, v in out.items() if k.endswith("Kb_data")})


if __name__ == "__main__":
    main()
