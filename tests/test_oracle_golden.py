"""Pins the CPU oracle (oracle/florence_oracle.c) to fixtures produced by running the reference itself
(tests/golden/make_golden.py): the reference's numpy twins of the LL formulas and its Fastor-free native code.
CPU only.  Tolerances: integer outputs bit-exact; fp64 within 1e-12 relative (rounding order only)."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix

from oracle import oracle as orc

MECH = ["LinearElastic", "NeoHookean", "MooneyRivlin", "NearlyIncompressibleMooneyRivlin", "ExplicitMooneyRivlin"]
ELEC = ["IsotropicElectroMechanics_101", "IsotropicElectroMechanics_105", "IsotropicElectroMechanics_108"]


def prm_for(name):
    mu, lamb, eps = 4.0e5, 2.0e6, 4.0 * 8.8541e-12
    if name in ("LinearElastic", "NeoHookean"):
        return orc.params(mu=mu, lamb=lamb)
    if name in ("MooneyRivlin", "ExplicitMooneyRivlin"):
        return orc.params(mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb)
    if name == "NearlyIncompressibleMooneyRivlin":
        # NearlyIncompressibleMooneyRivlin.py:30-34
        alpha = mu / 2.
        beta = (mu - 2. * alpha) / 3. / np.sqrt(3.)
        kappa = lamb + 4.0 / 3.0 * alpha + 2.0 * np.sqrt(3.0) * beta
        return orc.params(mu1=alpha, mu2=beta, mu3=kappa)
    if name.endswith("101"):
        return orc.params(mu=mu, lamb=lamb, eps_1=eps)
    if name.endswith("105"):
        return orc.params(mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb, eps_1=eps, eps_2=2.5 * eps)
    return orc.params(mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb, eps_2=eps)


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("name", MECH + ELEC)
def test_material_point_vs_reference_python_twin(golden, name, ndim):
    g = golden.materials
    Fs, Es = g["mat_F_%dd" % ndim], g["mat_E_%dd" % ndim]
    num = orc.MATERIAL_NUMBERS[name]
    for k in range(Fs.shape[0]):
        D, S, H = orc.material_point(num, Fs[k], Es[k], prm_for(name), want_hessian=name != "ExplicitMooneyRivlin")
        Sref = g["mat_%s_%dd_stress" % (name, ndim)][k]
        assert np.abs(S - Sref).max() <= 1e-12 * max(np.abs(Sref).max(), 4e5), (name, k)
        if name in ELEC:
            assert relerr(D, g["mat_%s_%dd_D" % (name, ndim)][k]) < 1e-13
        if name != "ExplicitMooneyRivlin":
            Href = g["mat_%s_%dd_hessian" % (name, ndim)][k]
            # blocks of the electro Hessian differ by ~20 orders of magnitude: compare block-wise
            hs = 6 if ndim == 3 else 3
            assert relerr(H[:hs, :hs], Href[:hs, :hs]) < 1e-12, (name, k)
            if name in ELEC:
                assert relerr(H[:hs, hs:], Href[:hs, hs:]) < 1e-12
                assert relerr(H[hs:, :hs], Href[hs:, :hs]) < 1e-12
                assert relerr(H[hs:, hs:], Href[hs:, hs:]) < 1e-12


def test_identity_F_gives_zero_stress():
    # intent of Florence/Utils/debug.py:60-81
    for name in ["NeoHookean", "MooneyRivlin", "NearlyIncompressibleMooneyRivlin"]:
        _, S, _ = orc.material_point(orc.MATERIAL_NUMBERS[name], np.eye(3), None, prm_for(name))
        assert np.abs(S).max() < 1e-9 * 4e5


def _case(g, key):
    names = ("points", "elements", "Eulerx", "Jm", "AllGauss", "K_data", "K_indices", "K_indptr", "T", "update", "prm", "sp_indices", "sp_indptr")
    d = {n: g[key + "_" + n] for n in names}
    d["Eulerp"] = g[key + "_Eulerp"] if key + "_Eulerp" in g.files else None
    if int(d["update"]) == 0:
        # linear analyses never move the geometry (Eulerx == mesh.points throughout, FEMSolver.py:319-324): the Python
        # path then takes SpatialGradient = MaterialGradient while the LL path still builds it from Eulerx
        # (_KinematicMeasures_.h:101-102); they agree only for the Eulerx the reference actually passes.
        d["Eulerx"] = d["points"]
    d["sp_dl"] = g[key + "_sp_dl"] if key + "_sp_dl" in g.files else None
    d["sp_dg"] = g[key + "_sp_dg"] if key + "_sp_dg" in g.files else None
    return d


def _asm_cases():
    import os
    out = []
    for f in ("assembly", "assembly_hi", "assembly_tet3"):
        d = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_%s.npz" % f))
        out += [(f, str(s)) for s in d["asm_cases"]]
    return out


@pytest.mark.parametrize("fixture,key", _asm_cases())
def test_implicit_assembly_vs_reference_python_path(golden, fixture, key):
    c = _case(getattr(golden, fixture), key)
    matname = key.split("_", 3)[3]
    num = orc.MATERIAL_NUMBERS[matname]
    nnode, ndim = c["points"].shape
    electro = matname in ELEC
    nvar = ndim + (1 if electro else 0)
    H = orc.hessian_size(num, ndim)
    n = nvar * nnode
    Kref = csr_matrix((c["K_data"], c["K_indices"], c["K_indptr"]), shape=(n, n))
    # COO mode (recompute_sparsity_pattern=True)
    I, J, V, T = orc.assemble_implicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, H,
                                       int(c["update"]), c["prm"], num, mode="coo")
    K = csr_matrix((V, (I, J)), shape=(n, n))
    if electro:
        # mechanical / coupling / dielectric blocks live on different scales
        mech = np.arange(n) % nvar != ndim
        for ra in (mech, ~mech):
            for ca in (mech, ~mech):
                A, B = K[ra][:, ca], Kref[ra][:, ca]
                assert abs(A - B).max() <= 1e-11 * abs(B).max(), key
        assert np.abs(T - c["T"])[mech].max() <= 1e-11 * np.abs(c["T"][mech]).max()
        assert np.abs(T - c["T"])[~mech].max() <= 1e-11 * max(np.abs(c["T"][~mech]).max(), 1e-300)
    else:
        assert abs(K - Kref).max() <= 1e-12 * abs(Kref).max(), key
        assert np.abs(T - c["T"]).max() <= 1e-12 * max(np.abs(c["T"]).max(), 1.0)
    # sparsity pattern: bit-exact against the reference's compiled ComputeSparsityPattern
    pat = orc.sparsity_pattern(c["elements"], nnode, nvar)
    assert pat[0].dtype == np.int32 and pat[1].dtype == np.int32
    assert np.array_equal(pat[0], c["sp_indices"]) and np.array_equal(pat[1], c["sp_indptr"])
    if c["sp_dl"] is not None:
        assert np.array_equal(pat[2], c["sp_dl"]) and np.array_equal(pat[3], c["sp_dg"])
    # CSR slot-map mode and binary-search mode agree with COO mode summed by scipy
    V1, T1 = orc.assemble_implicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, H,
                                   int(c["update"]), c["prm"], num, mode="csr", pattern=pat)
    se, so = orc.element_sorter(c["elements"])
    V2, T2 = orc.assemble_implicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, H,
                                   int(c["update"]), c["prm"], num, mode="csr_search", pattern=(pat[0], pat[1], se, so))
    K1 = csr_matrix((V1, pat[0], pat[1]), shape=(n, n))
    assert np.array_equal(V1, V2)
    scale = np.abs(V1).max()
    Kd = K.copy(); Kd.sum_duplicates(); Kd.sort_indices()
    assert abs(K1 - Kd).max() <= 1e-13 * scale
    assert np.array_equal(T1, T) and np.array_equal(T2, T)


def test_laplacian_vs_reference_python_path(golden):
    g = golden.laplacian
    for key in [str(s) for s in g["lap_cases"]]:
        pts, els = g[key + "_points"], g[key + "_elements"]
        n = pts.shape[0]
        Kref = csr_matrix((g[key + "_K_data"], g[key + "_K_indices"], g[key + "_K_indptr"]), shape=(n, n))
        # the LL wrapper passes e = -material.e (_LowLevelAssemblyPerfectLaplacian_.pyx:81); the python path integrates
        # with H = material.Permittivity: sign checked here
        I, J, V = orc.assemble_laplacian(pts, els, g[key + "_Jm"], g[key + "_AllGauss"], -g[key + "_e"], True, mode="coo")
        K = csr_matrix((V, (I, J)), shape=(n, n))
        sgn = 1.0 if abs(K - Kref).max() < abs(K + Kref).max() else -1.0
        assert abs(K - sgn * Kref).max() <= 1e-12 * abs(Kref).max(), key
        pat = orc.sparsity_pattern(els, n, 1)
        V1 = orc.assemble_laplacian(pts, els, g[key + "_Jm"], g[key + "_AllGauss"], -g[key + "_e"], True, mode="csr", pattern=pat)
        K1 = csr_matrix((V1, pat[0], pat[1]), shape=(n, n))
        assert abs(K1 - K).max() <= 1e-13 * abs(Kref).max()


def test_consistent_and_lumped_mass_vs_reference_python_path(golden):
    """a18: the consistent mass of the reference's Python path (Assemble, analysis_type="dynamic", optimise=False; fixtures of
    tests/golden/make_golden.py gen_mass) pins __TotalConstantMassIntegrand__'s consistent branch, incl. the zero rows of the
    potential dof (nvar = 4) and cubic tetrahedra; the lumped vector is its row sum (_MassIntegrand_.h:352-366)."""
    g = golden.mass
    for key in [str(s) for s in g["mass_cases"]]:
        P, E = g[key + "_points"], g[key + "_elements"]
        nvar, rho = int(g[key + "_nvar"]), float(g[key + "_rho"])
        n = nvar * P.shape[0]
        Mref = csr_matrix((g[key + "_M_data"], g[key + "_M_indices"], g[key + "_M_indptr"]), shape=(n, n))
        I, J, V = orc.assemble_mass(P, E, g[key + "_Bases"], g[key + "_Jm"], g[key + "_AllGauss"], nvar, rho, "consistent")
        M = csr_matrix((V, (I, J)), shape=(n, n))
        assert abs(M - Mref).max() <= 1e-13 * abs(Mref).max(), key
        Ml = orc.assemble_mass(P, E, g[key + "_Bases"], g[key + "_Jm"], g[key + "_AllGauss"], nvar, rho, "lumped")
        assert np.abs(Ml - np.asarray(Mref.sum(axis=1)).ravel()).max() <= 1e-13 * np.abs(Ml).max(), key
        if nvar > P.shape[1]:
            assert np.abs(Ml[P.shape[1]::nvar]).max() == 0.0, key


def test_explicit_force_mass_and_trajectory_vs_reference_integrator(golden):
    g = golden.explicit
    pts, els, Jm, AG, Bases = g["exp_points"], g["exp_elements"], g["exp_Jm"], g["exp_AllGauss"], g["exp_Bases"]
    prm, rho, dt, nsteps = g["exp_prm"], float(g["exp_rho"]), float(g["exp_dt"]), int(g["exp_nsteps"])
    nnode = pts.shape[0]
    # lumped mass and first internal force
    M = orc.assemble_mass(pts, els, Bases, Jm, AG, 3, rho, "lumped")
    assert relerr(M, g["exp_M_lumped"]) < 1e-13
    T0 = orc.assemble_explicit(pts, els, pts, None, Jm, AG, 3, prm, 1)
    assert np.abs(T0 - g["exp_T0"]).max() < 1e-9  # undeformed: T ~ 0 (reference holds rounding noise as well)
    fixed = np.zeros(nnode * 3, bool)
    fixed[g["exp_columns_out"]] = True
    NF, AD = g["exp_neumann"], g["exp_applied_dirichlet"]
    snaps, Eulerx, T = orc.explicit_central_difference(
        lambda X: orc.assemble_explicit(pts, els, X, None, Jm, AG, 3, prm, 1), pts, M, lambda inc: NF[:, inc - 1] if inc > 0 else NF[:, 0],
        dt, nsteps, fixed, applied_dirichlet_of=lambda inc: AD[:, inc - 1])
    ref = g["exp_TotalDisp"]
    for k, inc in enumerate(range(2, nsteps)):
        U = snaps[k].reshape(nnode, 3)
        assert np.abs(U - ref[:, :, inc]).max() <= 1e-11 * np.abs(ref).max(), inc


def test_explicit_equals_implicit_traction(golden):
    """a12 and a5 integrate the same B^T sigma: T from the matrix-free path equals the reference's T of the implicit path, for every
    fixture whose geometry is updated (all but the linear ones): 2-D and 3-D, p = 1..3, mechanics and electro-mechanics."""
    cases = [(f, k) for f, k in _asm_cases() if "LinearElastic" not in k]
    assert len(cases) >= 20
    for fixture, key in cases:
        g = getattr(golden, fixture)
        form = 1 if key.split("_", 3)[3] in ELEC else 0
        nvar = g[key + "_points"].shape[1] + form
        c = _case(g, key)
        num = orc.MATERIAL_NUMBERS[key.split("_", 3)[3]]
        T = orc.assemble_explicit(c["points"], c["elements"], c["Eulerx"], c["Eulerp"], c["Jm"], c["AllGauss"], nvar, c["prm"], num, form)
        mech = np.arange(T.shape[0]) % nvar < c["points"].shape[1]
        assert np.abs(T - c["T"])[mech].max() <= 1e-11 * np.abs(c["T"][mech]).max(), key
        if form:
            assert np.abs(T - c["T"])[~mech].max() <= 1e-11 * np.abs(c["T"][~mech]).max(), key
