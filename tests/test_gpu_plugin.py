"""GPU: the reference-facing plug-in functions (florence_b200/assembly.py) with duck-typed Florence objects, against the
fixtures produced by the reference itself.  These read like the reference's own usage: f(fem_solver, function_space,
formulation, mesh, material, Eulerx, Eulerp)."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Obj(object):
    pass


def make_objects(g, key, matname, fields, recompute, squeeze=False):
    pts, els = g[key + "_points"], g[key + "_elements"].astype(np.uint64)
    ndim = pts.shape[1]
    fs, fo, me, so = Obj(), Obj(), Obj(), Obj()
    fs.Bases, fs.Jm, fs.AllGauss = g[key + "_Bases"], g[key + "_Jm"], g[key + "_AllGauss"]
    fo.ndim, fo.fields = ndim, fields
    fo.nvar = ndim + (1 if fields == "electro_mechanics" else 0)
    me.points, me.elements, me.nelem, me.nnode = pts, els, els.shape[0], pts.shape[0]
    me.ChangeType = lambda: None
    me.GetNumberOfNodes = lambda: pts.shape[0]
    so.recompute_sparsity_pattern, so.squeeze_sparsity_pattern = recompute, squeeze
    so.requires_geometry_update = bool(int(g[key + "_update"]))
    prm = g[key + "_prm"]
    mat = type(matname, (object,), {})()
    for k, v in zip(("mu", "mu1", "mu2", "mu3", "mue", "lamb", "eps_1", "eps_2", "eps_3", "eps_e"), prm):
        setattr(mat, k, float(v))
    mat.alpha, mat.beta, mat.kappa = float(prm[1]), float(prm[2]), float(prm[3])
    mat.mtype, mat.rho = matname, 1100.0
    return so, fs, fo, me, mat


@pytest.mark.parametrize("key,fields", [("asm_hex2_n2_NeoHookean", "mechanics"), ("asm_tet2_n2_LinearElastic", "mechanics"),
                                        ("asm_quad2_n3_NeoHookean", "mechanics"), ("asm_hex2_n1_IsotropicElectroMechanics_108", "electro_mechanics"),
                                        ("asm_hex2_n1_NearlyIncompressibleMooneyRivlin", "mechanics")])
def test_low_level_assembly_dispatch(key, fields):
    from florence_b200 import assembly
    g = np.load(os.path.join(GOLD, "golden_assembly.npz"))
    matname = key.split("_", 3)[3]
    Eulerx = g[key + "_Eulerx"] if int(g[key + "_update"]) else g[key + "_points"]
    Eulerp = g[key + "_Eulerp"] if key + "_Eulerp" in g.files else np.zeros(g[key + "_points"].shape[0])
    n = g[key + "_T"].shape[0]
    Kref = csr_matrix((g[key + "_K_data"], g[key + "_K_indices"], g[key + "_K_indptr"]), shape=(n, n))
    nvar = n // g[key + "_points"].shape[0]
    ndim = g[key + "_points"].shape[1]
    mech = np.arange(n) % nvar < ndim

    def check(K, T):
        for ra in (mech, ~mech):
            for ca in (mech, ~mech):
                if ra.any() and ca.any():
                    assert abs(K[ra][:, ca] - Kref[ra][:, ca]).max() <= 1e-10 * abs(Kref[ra][:, ca]).max()
        assert np.linalg.norm((T - g[key + "_T"])[mech]) <= 1e-11 * max(np.linalg.norm(g[key + "_T"][mech]), 1e-300)

    # default mode of the reference: COO triplets -> scipy
    so, fs, fo, me, mat = make_objects(g, key, matname, fields, recompute=True)
    K, T, F, M = assembly._LowLevelAssembly_(so, fs, fo, me, mat, Eulerx, Eulerp)
    assert F == [] and M == [] and K.shape == (n, n) and T.shape == (n,)
    check(K, T)
    # pre-computed pattern mode: fem_solver.indices/indptr as FEMSolver.ComputeSparsityFEM stores them
    so, fs, fo, me, mat = make_objects(g, key, matname, fields, recompute=False)
    so.indices, so.indptr, so.data_local_indices, so.data_global_indices = assembly.ComputeSparsityPattern(me, nvar, function_space=fs)
    assert np.array_equal(so.indices, g[key + "_sp_indices"]) and np.array_equal(so.indptr, g[key + "_sp_indptr"])
    K2, T2, _, _ = assembly._LowLevelAssembly_(so, fs, fo, me, mat, Eulerx, Eulerp)
    check(K2, T2)
    V, T3 = assembly._LowLevelAssembly_Par_(so, fs, fo, me, mat, Eulerx, Eulerp)
    assert V.shape == so.indices.shape
    # squeeze_sparsity_pattern=True: only (indices, indptr) are returned and the same V comes back (binary-search mode of the reference)
    sq = assembly.ComputeSparsityPattern(me, nvar, squeeze_sparsity_pattern=True, function_space=fs)
    assert len(sq) == 2 and np.array_equal(sq[0], so.indices)
    so.squeeze_sparsity_pattern = True
    V2, _ = assembly._LowLevelAssembly_Par_(so, fs, fo, me, mat, Eulerx, Eulerp)
    assert np.array_equal(V2, V)
    # explicit entry points
    if so.requires_geometry_update:
        Te, F, M = assembly._LowLevelAssemblyExplicit_(so, fs, fo, me, mat, Eulerx, Eulerp)
        assert np.linalg.norm((Te - g[key + "_T"])[mech]) <= 1e-11 * np.linalg.norm(g[key + "_T"][mech])
        Tp = assembly._LowLevelAssemblyExplicit_Par_(fs, fo, me, mat, Eulerx, Eulerp)
        assert np.array_equal(Tp, Te)
    assembly.clear_handles()


def test_unknown_material_and_laplacian_dispatch():
    from florence_b200 import assembly
    g = np.load(os.path.join(GOLD, "golden_assembly.npz"))
    key = "asm_hex1_n2_NeoHookean"
    so, fs, fo, me, mat = make_objects(g, key, "SomeOtherMaterial", "mechanics", recompute=True)
    with pytest.raises(NotImplementedError):
        assembly._LowLevelAssembly_(so, fs, fo, me, mat, g[key + "_Eulerx"], np.zeros(me.nnode))
    with pytest.raises(NotImplementedError):
        assembly._LowLevelAssemblyExplicit_(so, fs, fo, me, mat, g[key + "_Eulerx"], np.zeros(me.nnode))
    gl = np.load(os.path.join(GOLD, "golden_laplacian.npz"))
    for key in ("lap_hex4_n1", "lap_tet2_n2", "lap_quad2_n3"):
        fs, fo, me, so, mat = Obj(), Obj(), Obj(), Obj(), Obj()
        pts, els = gl[key + "_points"], gl[key + "_elements"].astype(np.uint64)
        fs.Jm, fs.AllGauss, fs.Bases = gl[key + "_Jm"], gl[key + "_AllGauss"], None
        fo.ndim, fo.nvar, fo.fields = pts.shape[1], 1, "electrostatics"
        me.points, me.elements, me.nelem, me.nnode = pts, els, els.shape[0], pts.shape[0]
        me.ChangeType = lambda: None
        me.GetNumberOfNodes = lambda: None
        so.recompute_sparsity_pattern, so.squeeze_sparsity_pattern = True, False
        mat.e = gl[key + "_e"]
        K, T = assembly._LowLevelAssemblyLaplacian_(so, fs, fo, me, mat, pts, np.zeros(me.nnode))
        Kref = csr_matrix((gl[key + "_K_data"], gl[key + "_K_indices"], gl[key + "_K_indptr"]), shape=K.shape)
        # the LL wrapper assembles with -material.e (.pyx:81); the python-path fixture carries the opposite sign convention
        assert min(abs(K - Kref).max(), abs(K + Kref).max()) <= 1e-10 * abs(Kref).max()
        assert T.shape == (me.nnode,) and not T.any()
        mat.e = np.eye(pts.shape[1] + 1)
        with pytest.raises(ValueError):
            assembly._LowLevelAssemblyLaplacian_(so, fs, fo, me, mat, pts, np.zeros(me.nnode))
    assembly.clear_handles()


def test_explicit_integrator_solver_signature_reproduces_reference_trajectory():
    """ExplicitStructuralDynamicIntegrator.Solver with the reference's argument list, fed with what FEMSolver.Solve hands it
    (recorded in the fixture), must reproduce the TotalDisp the reference's own integrator produced."""
    from florence_b200 import assembly
    from florence_b200.time_integrator import ExplicitStructuralDynamicIntegrator as ESDI
    g = np.load(os.path.join(GOLD, "golden_explicit.npz"))
    pts, els = g["exp_points"], g["exp_elements"].astype(np.uint64)
    fs, fo, me, so, bc, mat = Obj(), Obj(), Obj(), Obj(), Obj(), type("NeoHookean", (object,), {})()
    fs.Bases, fs.Jm, fs.AllGauss = g["exp_Bases"], g["exp_Jm"], g["exp_AllGauss"]
    fo.ndim, fo.nvar, fo.fields = 3, 3, "mechanics"
    me.points, me.elements, me.nelem = pts, els, els.shape[0]
    me.ChangeType = lambda: None
    nsteps = int(g["exp_nsteps"])
    so.number_of_load_increments, so.total_time, so.mass_type, so.save_frequency = nsteps, float(g["exp_dt"]) * nsteps, "lumped", 1
    so.include_physical_damping, so.is_mass_computed, so.recompute_sparsity_pattern = False, False, True
    bc.columns_out, bc.applied_dirichlet, bc.make_loading = g["exp_columns_out"], g["exp_applied_dirichlet"], "ramp"
    prm = g["exp_prm"]
    mat.mu, mat.lamb, mat.rho, mat.mtype = float(prm[0]), float(prm[5]), float(g["exp_rho"]), "NeoHookean"
    # AssembleExplicit's first call: T of the undeformed mesh and the lumped mass (Assembly.py:664-715)
    T0, _, M = assembly.AssembleExplicit(so, fs, fo, me, mat, pts.copy(), np.zeros(pts.shape[0]))
    assert so.is_mass_computed is True
    assert np.abs(M.ravel() - g["exp_M_lumped"]).max() <= 1e-12 * np.abs(g["exp_M_lumped"]).max()
    TotalDisp = np.zeros((pts.shape[0], 3, nsteps))
    Eulerx = pts.copy()
    out = ESDI.Solver([fs, fs], fo, None, T0, M, g["exp_neumann"], None, None, me, TotalDisp, Eulerx, np.zeros(pts.shape[0]), mat, bc, so)
    ref = g["exp_TotalDisp"]
    assert out.shape == ref.shape
    for inc in range(2, nsteps):
        assert np.abs(out[:, :, inc] - ref[:, :, inc]).max() <= 1e-10 * np.abs(ref).max(), inc
    assembly.clear_handles()
