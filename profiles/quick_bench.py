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
        if num == 10:
            h.set_option(2, 0)
            V2, T2 = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
            t = med(lambda: h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V2, T2)))
            print("   block-wide kernel (option 2 = 0): elem=%.3f ms; identical K: %s, identical T: %s" % (t[0], torch.equal(V, V2), torch.equal(T, T2)))
            h.set_option(2, 1); del V2, T2
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
if len(sys.argv) > 3 and int(sys.argv[3]):
    n4 = int(sys.argv[3])
    # config 4: IsotropicElectroMechanics_108 p=3 hexes, one Newton-step assembly (K 9x9 Hessian + geometric, T), CSR
    pts, els = flmesh.box_hex_mesh(n4, n4, n4, p=3, device=dev)
    B, Jm, AG = flmesh.tables("hex", 3)
    x = flmesh.perturbed_state(pts, 1.0 / (3 * n4), 0.02, seed=1)
    phi = 9e3 * pts[:, 2] + 10.0 * (2 * torch.rand(pts.shape[0], dtype=torch.float64, device=dev) - 1)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    nnz = h.build_pattern(4); h.set_timing(True)
    mu = 5e4; lamb = 2 * mu * 0.4 / (1 - 0.8)
    mat = backend.make_material(8, 1200.0, mu1=mu, mu2=mu, lamb=lamb, eps_2=4 * 8.8541e-12)
    V, T = h.assemble_implicit(x, phi, mat, 1, True, mode="csr")
    fl = 64 * (2 * 81 * 256 + 2 * 9 * 65536 + 2 * 65536 + 25 * 4096)
    for opt in (1, 0):
        h.set_option(1, opt)
        t = med(lambda: h.assemble_implicit(x, phi, mat, 1, True, mode="csr", out=(V, T)), reps=3)
        print("hex64 EM_108 implicit csr (dmma=%d): nelem=%d nnz=%d elem=%.3f ms gather=%.3f ms -> %.3f Melem/s, %.2f TF/s (reference-count flops)" %
              (opt, els.shape[0], nnz, t[0], t[1], els.shape[0] / t.sum() / 1e3, fl * els.shape[0] / t[0] / 1e9))
    h.close(); del V, T
    # hex27 NeoHookean implicit (nonlinear Newton assembly at p=2)
    pts, els = flmesh.box_hex_mesh(32, 32, 32, p=2, device=dev)
    B, Jm, AG = flmesh.tables("hex", 2)
    x = flmesh.perturbed_state(pts, 1.0 / 64, 0.02, seed=1)
    h = backend.AssemblyHandle(pts, els, Jm, AG, B, device=dev)
    h.build_pattern(3); h.set_timing(True)
    mat = backend.make_material(1, 1100.0, mu=4e5, lamb=2e6)
    V, T = h.assemble_implicit(x, None, mat, 0, True, mode="csr")
    for opt in (1, 0):
        h.set_option(1, opt)
        t = med(lambda: h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T)), reps=3)
        print("hex27 NeoHookean implicit csr (dmma=%d): nelem=%d elem=%.3f ms gather=%.3f ms -> %.2f Melem/s" % (opt, els.shape[0], t[0], t[1], els.shape[0] / t.sum() / 1e3))
    h.close()
