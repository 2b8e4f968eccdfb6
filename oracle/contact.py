"""CPU restatement (numpy) of the explicit solver's rigid-plane penalty contact.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu legs, never by the product.

Follows ExplicitPenaltyContactFormulation.AssembleTractions
(Florence/VariationalPrinciple/ExplicitPenaltyContactFormulation.py:145-184) and its use in the time loop
(Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:190-197).
Pinned against the reference's own class by tests/golden/make_golden_contact.py -> tests/golden/golden_contact.npz.
"""
import numpy as np


def assemble_tractions(boundary_surface, Eulerx, plane_normal, distance, kappa, contact_gap_tolerance=1e-6):
    """:145-184; boundary_surface = mesh.faces (3-D) or mesh.edges (2-D).  Returns the flat (nnode*ndim,) contact force."""
    normal = np.asarray(plane_normal, dtype=np.float64).ravel()
    nnode, ndim = Eulerx.shape
    surf = np.unique(boundary_surface)                               # :161
    x = Eulerx[surf, :]
    gNx = x.dot(normal) + distance                                   # :163
    T = np.zeros((nnode, ndim))
    hit = gNx < contact_gap_tolerance                                # :166
    if not hit.any():
        return T.ravel()
    T[surf[hit].astype(np.int64), :] = kappa * np.outer(gNx[hit], normal)   # :180-182
    return T.ravel()
