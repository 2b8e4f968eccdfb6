"""Dirichlet reduction at config-2 size: device kernel time vs the reference's scipy slicing on the host."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, boundary, mesh as flmesh
from oracle import dirichlet as od
n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
host = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
pts, els = flmesh.box_tet_mesh(n, n, n, p=2, device=dev)
B, Jm, AG = flmesh.tables("tet", 2)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
nnz = h.build_pattern(3)
x = flmesh.perturbed_state(pts, 1.0 / n, 1e-3 * n, seed=1)
V, T = h.assemble_implicit(x, None, backend.make_material(10, 0.0, mu=1e5, lamb=1.5e5), 0, True, mode="csr")
flags = np.full((pts.shape[0], 3), np.nan)
z = pts[:, 2].cpu().numpy()
flags[np.isclose(z, 0.0)] = 0.0
flags[np.isclose(z, z.max()), 2] = 0.01
t0 = time.perf_counter()
bc = boundary.DeviceBoundaryCondition.from_flags(h, flags)
torch.cuda.synchronize()
print("build (maps + reduced pattern + export): %.1f ms; n_in=%d nnz_b=%d of nnz=%d" % (1e3 * (time.perf_counter() - t0), bc.n_in, bc.nnz_b, nnz))
F = T.clone()
Vb = torch.empty(bc.nnz_b, dtype=torch.float64, device=dev); Fb = torch.empty(bc.n_in, dtype=torch.float64, device=dev)
def run():
    h.dirichlet_apply(V, F, bc.applied_dirichlet, 0.5, out=(Vb, Fb))
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(10):
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print("device apply (K_b values + Dirichlet forces + F_b): %.3f ms -> %.0f GB/s (reads V, writes V_b)" % (ms, (nnz + bc.nnz_b) * 8 / ms / 1e6))
if host:
    indices, indptr = h.sparsity_pattern(3)
    N = 3 * pts.shape[0]
    K = od.full_csr(V.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy(), N)
    cols_out = bc.columns_out.cpu().numpy(); cols_in = od.columns_in(N, cols_out)
    Fh = T.cpu().numpy().copy()[:, None]
    t0 = time.perf_counter()
    Kb, Fbh, Fm = od.apply_dirichlet_get_reduced_matrices(K, Fh, bc.applied_dirichlet.cpu().numpy(), cols_in, cols_out, 0.5)
    th = time.perf_counter() - t0
    F2 = T.clone(); h.dirichlet_apply(V, F2, bc.applied_dirichlet, 0.5, out=(Vb, Fb)); torch.cuda.synchronize()
    print("host scipy (reference code path): %.1f ms; identical values: %s, identical F: %s" %
          (1e3 * th, np.array_equal(Kb.data, Vb.cpu().numpy()), np.array_equal(Fm[:, 0], F2.cpu().numpy())))
