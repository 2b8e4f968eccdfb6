"""Golden vectors of the rigid-plane penalty contact, produced by the REFERENCE's own class
(Florence/VariationalPrinciple/ExplicitPenaltyContactFormulation.py:145-184).

Run in the build container only:  python tests/golden/make_golden_contact.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _load_reference  # noqa: E402

Florence = _load_reference.load()
from Florence.VariationalPrinciple.ExplicitPenaltyContactFormulation import ExplicitPenaltyContactFormulation  # noqa: E402

from florence_b200 import mesh as flmesh  # noqa: E402
from oracle import contact as oc  # noqa: E402


class _Mesh(object):
    pass


def main():
    rng = np.random.default_rng(11)
    out = {}
    for tag, kind, p, n, ndim in (("hex27", "hex", 2, 3, 3), ("quad4", "quad", 1, 6, 2)):
        if kind == "hex":
            pts, els = flmesh.box_hex_mesh(n, n, n, p=p)
        else:
            pts, els = flmesh.rect_quad_mesh(n, n, p=p)
        pts, els = pts.numpy(), els.numpy()
        surf = flmesh.boundary_nodes(pts)
        m = _Mesh()
        m.points = pts
        # AssembleTractions only takes np.unique of the boundary connectivity: any array listing the surface nodes serves
        if ndim == 3:
            m.faces = surf.reshape(-1, 1)
        else:
            m.edges = surf.reshape(-1, 1)
        normal = np.array([0.0, 0.6, 0.8]) if ndim == 3 else np.array([0.6, 0.8])
        x = pts + 0.05 * rng.uniform(-1, 1, pts.shape)
        L = -0.35
        cf = ExplicitPenaltyContactFormulation(None, normal, L, 1e7, contact_gap_tolerance=1e-6)
        Tc = cf.AssembleTractions(m, None, x)
        To = oc.assemble_tractions(surf, x, normal, L, 1e7, 1e-6)
        assert np.array_equal(Tc.ravel(), To)
        far = ExplicitPenaltyContactFormulation(None, normal, 50.0, 1e7).AssembleTractions(m, None, x)      # no contact at all
        out.update({tag + "_points": pts, tag + "_elements": els.astype(np.int64), tag + "_surface_nodes": surf, tag + "_Eulerx": x,
                    tag + "_normal": normal, tag + "_distance": np.array(L), tag + "_kappa": np.array(1e7), tag + "_tol": np.array(1e-6),
                    tag + "_T_contact": Tc.ravel(), tag + "_T_far": np.asarray(far).ravel()})
        print(tag, "contact nodes:", int((np.abs(Tc.reshape(-1, ndim)).sum(1) > 0).sum()), "of", surf.shape[0], "surface nodes")
    np.savez_compressed(os.path.join(HERE, "golden_contact.npz"), **out)


if __name__ == "__main__":
    main()
