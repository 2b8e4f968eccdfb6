"""Poisson stiffness (a15, _LowLevelAssemblyPerfectLaplacian_) timing: p=4 hexahedra (config 1's element) on a larger mesh, and p=2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, mesh as flmesh
dev = torch.device("cuda:0")
for p, n in ((4, 16), (2, 48), (1, 128)):
    pts, els = flmesh.box_hex_mesh(n, n, n, p=p, device=dev)
    B, Jm, AG = flmesh.tables("hex", p)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    nnz = h.build_pattern(1); h.set_timing(True)
    e = -2.35 * np.eye(3)
    V = h.assemble_laplacian(e, True, mode="csr")
    ts = []
    for _ in range(5):
        h.assemble_laplacian(e, True, mode="csr"); ts.append(h.get_timing())
    t = np.median(np.array(ts), axis=0)
    npe = els.shape[1]
    print("hex p=%d Poisson csr: nelem=%d npe=%d nnz=%d elem=%.3f ms gather=%.3f ms -> %.2f Melem/s; K_e write %.0f GB/s" %
          (p, els.shape[0], npe, nnz, t[0], t[1], els.shape[0] / (t[0] + t[1]) / 1e3, els.shape[0] * npe * npe * 8 / t[0] / 1e6))
    h.close()
