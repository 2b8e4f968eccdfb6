"""CSR value reduction, row-buffer kernels (fl_set_option 3 = 0) against the register-resident gather (= 1) in element order, for the
shapes the register gather is instantiated for.  NeoHookean / electro-mechanics 108, meshes of 1-3 M elements."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, mesh as flmesh
dev = torch.device("cuda:0")
cases = [("tri", 1, 1200, 2), ("quad", 1, 1400, 2), ("tet", 1, 60, 4), ("tet", 1, 70, 3), ("hex", 1, 110, 3), ("tet", 2, 45, 3), ("tri", 2, 900, 2), ("quad", 2, 700, 2), ("tet", 2, 30, 4), ("hex", 1, 80, 4), ("hex", 2, 30, 4)]
for kind, p, n, nvar in cases:
    pts, els = flmesh.make_mesh(kind, n, p, device=dev)
    B, Jm, AG = flmesh.tables(kind, p)
    ndim = pts.shape[1]
    x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.02, seed=1)
    if nvar == ndim:
        mat, form, xp = backend.make_material(1, 1.0, mu=3.0, lamb=7.0), 0, None
    else:
        mat, form = backend.make_material(8, 1.0, mu1=2.0, mu2=1.5, lamb=6.0, eps_1=3.0, eps_2=2.0), 1
        xp = 0.1 * torch.sin(5.0 * pts.sum(1))
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    nnz = h.build_pattern(nvar)
    V = torch.empty(nnz, dtype=torch.float64, device=dev); T = torch.empty(pts.shape[0] * nvar, dtype=torch.float64, device=dev)
    res = {}
    for opt in (0, 1, 2):
        # 0 / 1: element order with the row-buffer kernels / the register gather; 2: K_e along the Morton curve + register gather
        os.environ.pop("FL_CURVE_ALL", None)
        if opt == 2:
            os.environ["FL_CURVE_ALL"] = "1"
        h.set_option(4, 1 if opt == 2 else 0)
        h.set_option(3, min(opt, 1))
        h.assemble_implicit(x, xp, mat, form, True, mode="csr", out=(V, T)); torch.cuda.synchronize()
        h.set_timing(True)
        ts = []
        for _ in range(5):
            h.assemble_implicit(x, xp, mat, form, True, mode="csr", out=(V, T)); ts.append(h.get_timing())
        h.set_timing(False)
        res[opt] = (np.median(np.array(ts[1:]), axis=0), V.clone())
    print("%-4s p=%d nvar=%d  %8d elements  element kernel %.3f ms (curve order %.3f)  reduction: row buffer %.3f ms, register gather %.3f ms, "
          "curve order %.3f ms  identical: %s" % (kind, p, nvar, els.shape[0], res[0][0][0], res[2][0][0], res[0][0][1], res[1][0][1], res[2][0][1],
                                                   torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][1], res[2][1])))
    h.close(); del V, T, h
    torch.cuda.empty_cache()
