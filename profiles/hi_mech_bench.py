import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, mesh as flmesh
n = 24
dev = torch.device("cuda:0")
pts, els = flmesh.box_hex_mesh(n, n, n, p=3, device=dev)
B, Jm, AG = flmesh.tables("hex", 3)
x = flmesh.perturbed_state(pts, 1.0 / (3 * n), 0.02, seed=11)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
mat = backend.make_material(2, 1200.0, mu1=5e4, mu2=5e4, lamb=2e5)
nnz = h.build_pattern(3); h.set_timing(True)
for o5 in (0, 1):
    h.set_option(5, o5)
    V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
    ts = []
    for _ in range(6):
        h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T)); ts.append(h.get_timing())
    t = np.median(np.array(ts[1:]), axis=0)
    print("hex64 MooneyRivlin 24^3 option5=%d: elements %.3f ms reduction %.3f ms" % (o5, t[0], t[1]))
