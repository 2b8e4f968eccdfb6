"""hex8 MooneyRivlin matrix-free internal force (config 5a kernel), 200^3 elements; FL_B200_LIB selects an alternative build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, mesh as flmesh
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
pts, els = flmesh.box_hex_mesh(n, n, n, p=1, device=dev)
B, Jm, AG = flmesh.tables("hex", 1)
x = flmesh.perturbed_state(pts, 1.0 / n, 0.02, seed=1)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
mat = backend.make_material(2, 1100.0, mu1=2e5, mu2=2e5, lamb=2e6)
T = torch.empty(pts.shape[0] * 3, dtype=torch.float64, device=dev)
h.set_timing(True)
ts = []
for _ in range(8):
    h.assemble_explicit(x, None, mat, 0, out=T); ts.append(h.get_timing())
t = np.median(np.array(ts[2:]), axis=0)
print("%-24s force kernel %.3f ms  nodal reduction %.3f ms  checksum %.12e" % (os.environ.get("FL_B200_LIB", "default").split("/")[-1], t[0], t[2], float(T.abs().sum())))
