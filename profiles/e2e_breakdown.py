"""Where the end-to-end (host in / host out) time of one implicit CSR assembly goes: handle lookup, H2D, kernels, D2H."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import assembly, backend, mesh as flmesh
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
dev = torch.device("cuda:0")
pts, els = flmesh.box_tet_mesh(n, n, n, p=2, device=dev)
Bases, Jm, AG = flmesh.tables("tet", 2)
x = flmesh.perturbed_state(pts, 1.0 / n, 1e-3 * n, seed=1)
x_host = x.cpu().numpy(); pts_host = pts.cpu().numpy(); els_host = els.cpu().numpy().astype(np.uint64)
so, fs, fo, me = bench.reference_objects(pts_host, els_host, Bases, Jm, AG, 3, 3, recompute=False)
material = bench._Obj(); material.mu, material.lamb, material.rho, material.mtype = 1e5, 1.5e5, 1.0, "LinearElastic"
func = assembly._LowLevelAssemblyDF__LinearElastic_
def T(f, reps=5):
    torch.cuda.synchronize(); out = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); out.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(out))
for _ in range(3):
    Vh, Th = func(so, fs, fo, me, material, x_host, None)
print("full call            %.2f ms" % T(lambda: func(so, fs, fo, me, material, x_host, None)))
print("get_handle           %.2f ms" % T(lambda: assembly.get_handle(me, fs)))
h = assembly.get_handle(me, fs)
print("state H2D            %.2f ms" % T(lambda: assembly._state_to_device(h, x_host, None)))
mat = assembly._material_struct(material, "LinearElastic")
xd, _ = assembly._state_to_device(h, x_host, None)
print("device assemble      %.2f ms" % T(lambda: h.assemble_implicit(xd, None, mat, 0, True, mode="csr")))
V, Tt = h.assemble_implicit(xd, None, mat, 0, True, mode="csr")
print("V D2H (%.2f GB)       %.2f ms" % (V.numel() * 8 / 1e9, T(lambda: assembly._to_host(V, "V"))))
print("T D2H                %.2f ms" % T(lambda: assembly._to_host(Tt, "T")))
buf = torch.empty(V.numel(), dtype=torch.float64, pin_memory=True)
ms = T(lambda: buf.copy_(V, non_blocking=True))
print("raw pinned D2H       %.2f ms  -> %.1f GB/s" % (ms, V.numel() * 8 / ms / 1e6))
half = V.numel() // 2
s2 = torch.cuda.Stream()
def two():
    buf[:half].copy_(V[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        buf[half:].copy_(V[half:], non_blocking=True)
ms = T(two)
print("2-stream pinned D2H  %.2f ms  -> %.1f GB/s" % (ms, V.numel() * 8 / ms / 1e6))
