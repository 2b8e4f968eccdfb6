"""GPU: the drop-in for ExplicitStructuralDynamicIntegrator.Solver against runs of the reference's own integrator
(tests/golden/make_golden.py explicit_rules): save_frequency, the growth blow-up test, break_at_increment and the trimming of
TotalDisp (ExplicitStructuralDynamicIntegrator.py:165-251), with the loop running in device-resident chunks."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Obj(object):
    pass


def _objects(g, tag):
    mesh, fs, form, bc, so, mat = (Obj() for _ in range(6))
    mesh.points, mesh.elements, mesh.faces = g["rule_points"], g["rule_elements"].astype(np.uint64), g["rule_faces"]
    mesh.nelem, mesh.nnode = mesh.elements.shape[0], mesh.points.shape[0]
    mesh.ChangeType = lambda: None
    fs.Jm, fs.AllGauss, fs.Bases = g["rule_Jm"], g["rule_AllGauss"], g["rule_Bases"]
    form.fields, form.ndim, form.nvar = "mechanics", 3, 3
    bc.columns_out, bc.applied_dirichlet, bc.make_loading = g["rule_columns_out"], g["rule_applied_dirichlet"], "ramp"
    bc.has_step_wise_dirichlet_loading = bc.has_step_wise_neumann_loading = False
    n = int(g["rule_%s_nsteps" % tag])
    so.number_of_load_increments, so.total_time = n, float(g["rule_%s_dt" % tag]) * n
    so.mass_type, so.include_physical_damping, so.has_contact = "lumped", False, False
    so.save_frequency = 3 if tag == "sf3" else 1
    so.break_at_increment = 9 if tag == "brk" else -1
    prm = g["rule_prm"]
    mat.mtype, mat.mu, mat.lamb, mat.rho = "NeoHookean", float(prm[0]), float(prm[5]), 1100.0
    return mesh, fs, form, bc, so, mat


@pytest.mark.parametrize("tag", ["sf3", "blow", "brk"])
def test_solver_dropin_follows_the_reference_save_and_termination_rules(tag):
    from florence_b200 import assembly, time_integrator
    g = np.load(os.path.join(GOLD, "golden_explicit_rules.npz"))
    mesh, fs, form, bc, so, mat = _objects(g, tag)
    pre = "rule_%s_" % tag
    ref = g[pre + "TotalDisp"]
    TotalDisp = np.zeros(tuple(g[pre + "TotalDisp_shape_in"]))
    Eulerx = mesh.points.copy()
    out = time_integrator.ExplicitStructuralDynamicIntegrator.Solver(
        [fs, fs], form, None, g[pre + "T0"], g[pre + "M"], g[pre + "neumann"], None, None, mesh, TotalDisp, Eulerx, np.zeros(mesh.nnode),
        mat, bc, so)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    assert so.number_of_load_increments == int(g[pre + "number_of_load_increments"])
    scale = np.abs(ref).max()
    if tag == "blow":
        # an unstable run amplifies rounding differences by the growth factor of every step: frames before the blow-up agree to
        # the digits that survive, the cut itself (shape, number_of_load_increments) is exact
        for k in range(ref.shape[2]):
            assert np.abs(out[:, :, k] - ref[:, :, k]).max() <= 1e-7 * max(np.abs(ref[:, :, k]).max(), 1e-300), k
    else:
        assert np.abs(out - ref).max() <= 1e-10 * scale
    assembly.clear_handles()


def test_growth_test_in_a_chunk_equals_the_step_by_step_test():
    """The blow-up test runs on the device after every update, also inside a multi-increment chunk: status bits and the increment
    of the first detection equal what single increments + the reference's formula on the host give."""
    from florence_b200 import backend, time_integrator
    g = np.load(os.path.join(GOLD, "golden_explicit_rules.npz"))
    pts, els = g["rule_points"], g["rule_elements"]
    h = backend.AssemblyHandle(pts, els, g["rule_Jm"], g["rule_AllGauss"], g["rule_Bases"])
    prm = g["rule_prm"]
    mat = backend.make_material(1, 1100.0, mu=float(prm[0]), lamb=float(prm[5]))
    dt = float(g["rule_blow_dt"])
    nn = pts.shape[0]
    fixed = np.zeros(nn * 3, np.uint8); fixed[g["rule_columns_out"]] = 1
    fext = torch.as_tensor(g["rule_blow_neumann"].ravel(), device=h.device)

    def fresh():
        integ = time_integrator.ExplicitStructuralDynamicIntegrator(h, mat, rho=1100.0)
        integ.initialise(pts, None, fixed, dt)
        return integ
    a = fresh()
    first, U0 = None, a.U0.clone()
    for inc in range(2, 30):
        prev = a.U0.clone()
        st = a.step(1, inc, fext, 0.0, 1.0 / 40)
        U = a.U0
        tol = 1e200 if inc < 5 else 10.0
        host = bool(torch.isnan(U).any()) or abs(float(U.max()) / (float(prev.max()) + 1e-14)) > tol
        assert bool(st) == host, inc
        if st:
            first = inc
            assert a.last_status[1] == inc
            break
    assert first is not None and first >= 5
    b = fresh()
    st = b.step(28, 2, fext, 0.0, 1.0 / 40)
    assert st & time_integrator.GROWTH_BIT and b.last_status[1] == first
    h.close()
