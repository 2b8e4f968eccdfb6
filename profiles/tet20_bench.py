"""tet20 (cubic tetrahedra, the reference's 14-point rule): generic column-owner kernel vs the DMMA parent-space pair.

The package's own table generator stops at p = 2 for simplices, so the reference-generated tables of tests/golden/
golden_assembly_tet3.npz are used, on a conforming tet20 mesh built here from the package's tet4 connectivity.
    python profiles/tet20_bench.py [n]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
from florence_b200 import mesh as flmesh
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_assembly_tet3.npz"))
dev = torch.device("cuda:0")


def tet20_mesh(n, key):
    """Conforming tet20 mesh on n^3 cubes: the package's tet4 connectivity, the 20 nodes of every element placed at the
    barycentric positions of the reference's arrangement (read off the fixture's first element), duplicates merged."""
    P, E = g[key + "_points"], g[key + "_elements"]
    X = P[E[0]]
    lam = np.linalg.solve((X[1:4] - X[0]).T, (X - X[0]).T).T
    lam = np.c_[1.0 - lam.sum(1), lam]                                    # (20, 4)
    p4, e4 = flmesh.box_tet_mesh(n, n, n, p=1)
    p4, e4 = p4.numpy(), e4.numpy().astype(np.int64)
    allp = np.einsum("ak,ekd->ead", lam, p4[e4]).reshape(-1, 3)           # (nelem*20, 3)
    q = np.round(allp * (3.0 * n) * 1e6).astype(np.int64)
    _, first, inv = np.unique(q, axis=0, return_index=True, return_inverse=True)
    return allp[first], inv.reshape(-1, 20)


for key, num, form in (("asm_tet3_n1_NeoHookean", 1, 0), ("asm_tet3_n1_IsotropicElectroMechanics_108", 8, 1)):
    pts, els = tet20_mesh(n, key)
    rng = np.random.default_rng(5)
    x = pts + 0.02 / (3 * n) * rng.uniform(-1, 1, pts.shape)
    ph = torch.as_tensor(9e3 * pts[:, 2] + 10.0 * rng.uniform(-1, 1, pts.shape[0]), device=dev) if form else None
    h = backend.AssemblyHandle(torch.as_tensor(pts, device=dev), torch.as_tensor(els, device=dev), g[key + "_Jm"], g[key + "_AllGauss"],
                               g[key + "_Bases"], device=dev)
    prm = g[key + "_prm"]
    mat = backend.make_material(num, 0.0, mu=prm[0], mu1=prm[1], mu2=prm[2], mu3=prm[3], mue=prm[4], lamb=prm[5], eps_1=prm[6], eps_2=prm[7],
                                eps_3=prm[8], eps_e=prm[9])
    xt = torch.as_tensor(x, device=dev)
    nnz = h.build_pattern(3 + form); h.set_timing(True)
    res = {}
    for val in (0, 2):
        h.set_option(1, val)
        V, T = h.assemble_implicit(xt, ph, mat, form, True, mode="csr")
        ts = []
        for _ in range(7):
            h.assemble_implicit(xt, ph, mat, form, True, mode="csr", out=(V, T)); ts.append(h.get_timing())
        t = np.median(np.array(ts), axis=0)
        res[val] = (V.clone(), t)
        print("%s nelem=%d nnode=%d nnz=%d option1=%d: element %.3f ms, reduction %.3f ms, T %.3f ms -> %.2f M elements/s" %
              (key, els.shape[0], pts.shape[0], nnz, val, t[0], t[1], t[2], els.shape[0] / t.sum() / 1e3))
    d = (res[0][0] - res[2][0]).abs().max().item() / res[0][0].abs().max().item()
    print("   max |V_generic - V_dmma| / max|V| = %.2e" % d)
    h.close()
