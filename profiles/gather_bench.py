"""Config 2 (tet10 LinearElastic K(CSR)+T): register-resident CSR reduction (fl_set_option 3 = 1) against the shared-memory
row-buffer kernel (= 0).  FL_GATHER_U selects the number of visits in flight per lane."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, mesh as flmesh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
dev = torch.device("cuda:0")
pts, els = flmesh.box_tet_mesh(n, n, n, p=2, device=dev)
B, Jm, AG = flmesh.tables("tet", 2)
x = flmesh.perturbed_state(pts, 1.0 / n, 1e-3 * n, seed=0)
h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
mat = backend.make_material(10, 0.0, mu=1e5, lamb=1.5e5)
t0 = time.perf_counter(); nnz = h.build_pattern(3); torch.cuda.synchronize()
print("pattern + gather plan: %.3f s, nnz %d" % (time.perf_counter() - t0, nnz))
V = torch.empty(nnz, dtype=torch.float64, device=dev); T = torch.empty(pts.shape[0] * 3, dtype=torch.float64, device=dev)
out = {}
h.set_option(4, int(os.environ.get('FL_OPT4', '0')))
for opt in (0, 1):
    h.set_option(3, opt)
    h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T)); torch.cuda.synchronize()
    h.set_timing(True)
    ts = []
    for _ in range(8):
        h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T)); ts.append(h.get_timing())
    h.set_timing(False)
    t = np.median(np.array(ts[2:]), axis=0)
    out[opt] = (V.clone(), T.clone())
    print("U=%s option3=%d  kernels %.3f + %.3f + %.3f = %.3f ms  -> %.1f M elements/s" % (os.environ.get("FL_GATHER_U", "3"), opt, t[0], t[1], t[2], t.sum(), els.shape[0] / t.sum() / 1e3))
print("V bit-identical: %s   T bit-identical: %s" % (torch.equal(out[0][0], out[1][0]), torch.equal(out[0][1], out[1][1])))
