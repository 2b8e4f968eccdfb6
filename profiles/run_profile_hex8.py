import sys, os
sys.path.insert(0, "/root/repo")
import torch
from florence_b200 import backend, mesh as flmesh
dev = torch.device("cuda:0")
n = 160
pts, els = flmesh.box_hex_mesh(n, n, n, p=1, device=dev)
B, Jm, AG = flmesh.tables("hex", 1)
x = flmesh.perturbed_state(pts, 1.0 / n, 0.02, seed=1)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
mat = backend.make_material(2, 1100.0, mu1=2e5, mu2=2e5, lamb=2e6)
for _ in range(3):
    T = h.assemble_explicit(x, None, mat, 0)
torch.cuda.synchronize()
print("done")
