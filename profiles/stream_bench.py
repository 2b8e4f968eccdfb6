"""Config 2 (tet10 LinearElastic K(CSR)+T): the curve-ordered paths (fl_set_option 4: 1 two-pass, 2 concurrent, 3 sequential with
flags) against the element-order two-pass path (0) with either CSR reduction (option 3)."""
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
nnz = h.build_pattern(3)
V = torch.empty(nnz, dtype=torch.float64, device=dev); T = torch.empty(pts.shape[0] * 3, dtype=torch.float64, device=dev)
out = {}
only = os.environ.get("FL_ONLY")
for name, o3, o4 in (("element order, smem gather", 0, 0), ("element order, reg gather", 1, 0), ("curve order, two-pass", 0, 1), ("curve, flags, sequential", 0, 3), ("curve, concurrent", 0, 2)):
    if only and only not in name and name != "element order, smem gather":
        continue
    h.set_option(3, o3); h.set_option(4, o4)
    t0 = time.perf_counter()
    h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T)); torch.cuda.synchronize()
    first = time.perf_counter() - t0
    h.set_timing(True)
    ts = []; wall = []
    for _ in range(10):
        V.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T)); e1.record(); torch.cuda.synchronize()
        ts.append(h.get_timing()); wall.append(e0.elapsed_time(e1))
    h.set_timing(False)
    t = np.median(np.array(ts[2:]), axis=0)
    out[name] = (V.clone(), T.clone())
    print("%-28s first %.3f s  marks %.3f + %.3f + %.3f  step %.3f ms -> %.1f M elements/s" % (name, first, t[0], t[1], t[2], np.median(wall[2:]), els.shape[0] / np.median(wall[2:]) / 1e3))
ref = out["element order, smem gather"]
for name, (v, t) in out.items():
    print("%-28s V bit-identical: %s   T bit-identical: %s" % (name, torch.equal(v, ref[0]), torch.equal(t, ref[1])))
