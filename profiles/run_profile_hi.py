"""ncu driver for the high-order implicit kernel: hex64 IsotropicElectroMechanics_108 (config 4 shape) on a small mesh."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from florence_b200 import backend, mesh as flmesh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda:0")
pts, els = flmesh.box_hex_mesh(n, n, n, p=3, device=dev)
B, Jm, AG = flmesh.tables("hex", 3)
x = flmesh.perturbed_state(pts, 1.0 / (3 * n), 0.02, seed=1)
phi = 9e3 * pts[:, 2] + 10.0 * (2 * torch.rand(pts.shape[0], dtype=torch.float64, device=dev) - 1)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
mu = 5e4; lamb = 2 * mu * 0.4 / (1 - 0.8)
mat = backend.make_material(8, 1200.0, mu1=mu, mu2=mu, lamb=lamb, eps_2=4 * 8.8541e-12)
h.build_pattern(4)
for _ in range(1):
    V, T = h.assemble_implicit(x, phi, mat, 1, True, mode="csr")
torch.cuda.synchronize()
print("done hex64")
h.close()
# config 1's element: Poisson on hex125 (DMMA GEMM kernel with nvar = 1)
import numpy as np
n1 = max(4, n // 2)
pts, els = flmesh.box_hex_mesh(n1, n1, n1, p=4, device=dev)
B, Jm, AG = flmesh.tables("hex", 4)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
h.build_pattern(1)
for _ in range(2):
    V = h.assemble_laplacian(-2.35 * np.eye(3), True, mode="csr")
torch.cuda.synchronize()
print("done laplacian")
