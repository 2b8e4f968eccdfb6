"""Config-4 (hex64 EM_108, 24^3) element-kernel / reduction timing; FL_B200_LIB selects an alternative build of the library
(A/B timing of kernel schedules: profiles/r2_hex64_variants.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, mesh as flmesh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda:0")
pts, els = flmesh.box_hex_mesh(n, n, n, p=3, device=dev)
B, Jm, AG = flmesh.tables("hex", 3)
x = flmesh.perturbed_state(pts, 1.0 / (3 * n), 0.02, seed=11)
phi = 9e3 * pts[:, 2] + 10.0 * (2 * torch.rand(pts.shape[0], dtype=torch.float64, device=dev) - 1)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
mu = 5e4
mat = backend.make_material(8, 1200.0, mu1=mu, mu2=mu, lamb=2 * mu * 0.4 / (1 - 0.8), eps_2=4 * 8.8541e-12)
if os.environ.get('FL_OPT1'):
    h.set_option(1, int(os.environ['FL_OPT1']))   # 3: K_e scratch as dof-pair planes (round-2 layout before the per-row-node planes)
if os.environ.get('FL_OPT5'):
    h.set_option(5, int(os.environ['FL_OPT5']))   # 1: CSR reduction without the cross-node software pipeline
nnz = h.build_pattern(4)
V = torch.empty(nnz, dtype=torch.float64, device=dev); T = torch.empty(pts.shape[0] * 4, dtype=torch.float64, device=dev)
h.set_timing(True)
ts = []
for _ in range(6):
    h.assemble_implicit(x, phi, mat, 1, True, mode="csr", out=(V, T)); ts.append(h.get_timing())
t = np.median(np.array(ts[1:]), axis=0)
print("%-28s elements %.3f ms  reduction %.3f ms  total %.3f ms  checksum %.10e" % (os.environ.get("FL_B200_LIB", "default").split("/")[-1], t[0], t[1], t.sum(), float(V.abs().sum())))
