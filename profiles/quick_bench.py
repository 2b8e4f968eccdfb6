"""Quick kernel timing (CUDA events inside the library): tet10 implicit CSR assembly and hex27 explicit force."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from florence_b200 import backend, mesh as flmesh
n_tet = int(sys.argv[1]) if len(sys.argv) > 1 else 55
n_hex = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
def med(f, reps=7):
    out = []
    for _ in range(reps):
        f(); out.append(h.get_timing())
    return np.median(np.array(out), axis=0)
if n_tet:
    pts, els = flmesh.box_tet_mesh(n_tet, n_tet, n_tet, p=2, device=dev)
    B, Jm, AG = flmesh.tables("tet", 2)
    x = flmesh.perturbed_state(pts, 1.0 / n_tet, 1e-3 * n_tet, seed=1)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    h.build_pattern(3); h.set_timing(True)
    for name, num, prm in (("LinearElastic", 10, dict(mu=1e5, lamb=1.5e5)), ("NeoHookean", 1, dict(mu=1e5, lamb=1.5e5))):
        mat = backend.make_material(num, 0.0, **prm)
        V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
        t = med(lambda: h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T)))
        print("tet10 %s implicit csr: nelem=%d elem=%.3f ms gather=%.3f ms T=%.3f ms -> %.1f Melem/s (elem kernel %.1f Melem/s)" %
              (name, els.shape[0], t[0], t[1], t[2], els.shape[0] / t.sum() / 1e3, els.shape[0] / t[0] / 1e3))
    h.close(); del V, T
if n_hex:
    for p, n in ((2, n_hex), (1, 2 * n_hex)):
        pts, els = flmesh.box_hex_mesh(n, n, n, p=p, device=dev)
        B, Jm, AG = flmesh.tables("hex", p)
        x = flmesh.perturbed_state(pts, 1.0 / (p * n), 0.02, seed=1)
        h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
        h.set_timing(True)
        for name, num, prm in (("NeoHookean", 1, dict(mu=4e5, lamb=2e6)), ("MooneyRivlin", 2, dict(mu1=2e5, mu2=2e5, lamb=2e6))):
            mat = backend.make_material(num, 1100.0, **prm)
            for mma in (1, 0):
                h.set_option(0, mma)
                T = h.assemble_explicit(x, None, mat, 0)
                t = med(lambda: h.assemble_explicit(x, None, mat, 0, out=T))
                print("hex p=%d %s explicit (dmma=%d): nelem=%d elem=%.3f ms gather=%.3f ms -> %.1f Melem/s" % (p, name, mma, els.shape[0], t[0], t[2], els.shape[0] / (t[0] + t[2]) / 1e3))
        h.close()
