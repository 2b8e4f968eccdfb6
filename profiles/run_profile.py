"""Small driver for ncu captures: a few tet10 implicit CSR assemblies and hex27 explicit force evaluations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from florence_b200 import backend, mesh as flmesh

n_tet = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_hex = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
if n_tet > 0:
    pts, els = flmesh.box_tet_mesh(n_tet, n_tet, n_tet, p=2, device=dev)
    B, Jm, AG = flmesh.tables("tet", 2)
    x = flmesh.perturbed_state(pts, 1.0 / n_tet, 1e-3 * n_tet, seed=1)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    mat = backend.make_material(10, 0.0, mu=1e5, lamb=1.5e5)
    h.build_pattern(3)
    for _ in range(reps):
        V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
    torch.cuda.synchronize()
    h.close()
if n_hex > 0:
    pts, els = flmesh.box_hex_mesh(n_hex, n_hex, n_hex, p=2, device=dev)
    B, Jm, AG = flmesh.tables("hex", 2)
    x = flmesh.perturbed_state(pts, 0.5 / n_hex, 0.02, seed=1)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    mat = backend.make_material(1, 1100.0, mu=4e5, lamb=2e6)
    for _ in range(reps):
        T = h.assemble_explicit(x, None, mat, 0)
    torch.cuda.synchronize()
    h.close()
print("done")
