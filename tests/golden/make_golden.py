"""Generates the golden fixtures under tests/golden/ by RUNNING THE REFERENCE (romeric/florence, read-only
at /root/reference) in the build container.  Run once here: `python tests/golden/make_golden.py`.
The GPU box never runs this; it only reads the committed .npz files.

What is executed is the reference's own code (see _load_reference.py for what is stubbed):
  * FunctionSpace / QuadratureRule / Mesh (Python)              -> tables and meshes
  * MaterialLibrary/<Material>.py CauchyStress / Hessian (numpy) -> material-point vectors
  * DisplacementFormulation / DisplacementPotentialFormulation .GetLocalStiffness (numpy) through
    AssembleForm -> AssemblySmall + the reference's compiled SparseAssemblyNative / RHSAssemblyNative
    (Florence/FiniteElements/Assembly/Assembly.py:100-250)       -> global K (COO->CSR), T
  * compiled ComputeSparsityPattern (.pyx/.h)                    -> indices, indptr, slot maps
  * LaplacianFormulation (numpy path)                            -> Poisson K
  * VariationalPrinciple.GetLocalMass_Efficient / lumped mass, AssembleExplicit python path
  * ExplicitStructuralDynamicIntegrator (optimise=False)         -> a short explicit trajectory

Known differences of the Python path from the LL path that the oracle follows:
signed det J (meshes here have J>0) and makezero(|K_e|<1e-12) (absolute 1e-12 on entries of O(1e5)).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _load_reference  # noqa: E402

warnings.simplefilter("ignore")
Fl = _load_reference.load()
from Florence import (Mesh, FEMSolver, AssembleForm, BoundaryCondition, DisplacementFormulation,  # noqa: E402
                      DisplacementPotentialFormulation, LaplacianFormulation)
import Florence as F  # noqa: E402
from Florence.FiniteElements.LocalAssembly.KinematicMeasures import KinematicMeasures  # noqa: E402
from Florence.FiniteElements.Assembly.ComputeSparsityPattern import ComputeSparsityPattern  # noqa: E402


def make_mesh(etype, p, n):
    mesh = Mesh()
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    if etype in ("hex", "tet"):
        mesh.Parallelepiped(upper_right_front_point=(1.0, 0.8, 1.2), nx=nx, ny=ny, nz=nz, element_type=etype)
    else:
        mesh.Rectangle(upper_right_point=(1.0, 0.8), nx=n, ny=n, element_type=etype)
    if etype == "tet":
        # the reference's 6-tet split leaves a third of the tets negatively oriented; its Python path integrates with the
        # SIGNED det J (DisplacementFormulation.py:146) while the LL path takes fabs (_KinematicMeasures_.h:95-98).
        # The fixtures use positively oriented tets, where both agree.
        X = mesh.points[mesh.elements]
        vol = np.einsum("ei,ei->e", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0])
        # parent tet of the reference has det>0 for vol<0 ordering or vice versa: decide by the p=1 table
        form = DisplacementFormulation(mesh)
        Jm = form.function_spaces[0].Jm
        det = np.array([np.linalg.det(Jm[:, :, 0] @ mesh.points[e]) for e in mesh.elements])
        neg = det < 0
        mesh.elements[neg] = mesh.elements[neg][:, [1, 0, 2, 3]]
        mesh.GetBoundaryFacesTet(); mesh.GetBoundaryEdgesTet()
    if p > 1:
        mesh.GetHighOrderMesh(p=p)
    mesh.ChangeType()
    return mesh


def material_of(name, ndim):
    mu, lamb = 4.0e5, 2.0e6
    eps = 4.0 * 8.8541e-12
    kw = dict(rho=1100.0)
    if name == "LinearElastic":
        return F.LinearElastic(ndim, mu=mu, lamb=lamb, **kw), dict(mu=mu, lamb=lamb)
    if name == "NeoHookean":
        return F.NeoHookean(ndim, mu=mu, lamb=lamb, **kw), dict(mu=mu, lamb=lamb)
    if name in ("MooneyRivlin", "ExplicitMooneyRivlin"):
        m = getattr(F, name)(ndim, mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb, **kw)
        return m, dict(mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb)
    if name == "NearlyIncompressibleMooneyRivlin":
        m = F.NearlyIncompressibleMooneyRivlin(ndim, mu=mu, lamb=lamb, **kw)
        return m, dict(mu1=m.alpha, mu2=m.beta, mu3=m.kappa)
    if name == "IsotropicElectroMechanics_101":
        return F.IsotropicElectroMechanics_101(ndim, mu=mu, lamb=lamb, eps_1=eps, **kw), dict(mu=mu, lamb=lamb, eps_1=eps)
    if name == "IsotropicElectroMechanics_105":
        m = F.IsotropicElectroMechanics_105(ndim, mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb, eps_1=eps, eps_2=2.5 * eps, **kw)
        return m, dict(mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb, eps_1=eps, eps_2=2.5 * eps)
    if name in ("IsotropicElectroMechanics_108", "ExplicitIsotropicElectroMechanics_108"):
        m = getattr(F, name)(ndim, mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb, eps_2=eps, **kw)
        return m, dict(mu1=0.6 * mu, mu2=0.4 * mu, lamb=lamb, eps_2=eps)
    raise KeyError(name)


MECH = ["LinearElastic", "NeoHookean", "MooneyRivlin", "NearlyIncompressibleMooneyRivlin"]
ELEC = ["IsotropicElectroMechanics_101", "IsotropicElectroMechanics_105", "IsotropicElectroMechanics_108"]


def gen_tables(out):
    """Bases / Jm / AllGauss of the reference's FunctionSpace for the element types of the configs."""
    for etype, p in (("hex", 1), ("hex", 2), ("hex", 3), ("hex", 4), ("tet", 1), ("tet", 2), ("quad", 1), ("quad", 2), ("tri", 1), ("tri", 2)):
        mesh = make_mesh(etype, p, 1)
        form = DisplacementFormulation(mesh)
        for which, fs in (("", form.function_spaces[0]), ("_post", form.function_spaces[1])):
            out["tab_%s%d%s_Bases" % (etype, p, which)] = fs.Bases
            out["tab_%s%d%s_Jm" % (etype, p, which)] = fs.Jm
            out["tab_%s%d%s_AllGauss" % (etype, p, which)] = fs.AllGauss
        # node coordinates of the reference element (for our own generator's ordering check)
        out["tab_%s%d_points" % (etype, p)] = mesh.points
        out["tab_%s%d_elements" % (etype, p)] = mesh.elements.astype(np.int64)


def gen_materials(out):
    rng = np.random.default_rng(7)
    for ndim in (2, 3):
        Fs = np.eye(ndim)[None] + 0.15 * rng.uniform(-1, 1, (6, ndim, ndim))
        Fs[0] = np.eye(ndim)
        Es = 1.0e6 * rng.uniform(-1, 1, (6, ndim))
        out["mat_F_%dd" % ndim] = Fs
        out["mat_E_%dd" % ndim] = Es
        st = KinematicMeasures(Fs, "nonlinear")
        for name in MECH + ["ExplicitMooneyRivlin"]:
            m, prm = material_of(name, ndim)
            m.has_low_level_dispatcher = False
            S = np.array([m.CauchyStress(st, None, 0, g) for g in range(6)])
            out["mat_%s_%dd_stress" % (name, ndim)] = S
            if name != "ExplicitMooneyRivlin":
                H = np.array([m.Hessian(st, None, 0, g) for g in range(6)])
                out["mat_%s_%dd_hessian" % (name, ndim)] = H
        for name in ELEC:
            m, prm = material_of(name, ndim)
            m.has_low_level_dispatcher = False
            Ds, Ss, Hs = [], [], []
            for g in range(6):
                J, b = st["J"][g], st["b"][g]
                E = Es[g]
                # closed forms the LL kernels use for D (e.g. _IsotropicElectroMechanics_105_.h:52-53); the Python
                # twin's Newton iteration (LegendreTransform.GetElectricDisplacement) converges to the same D.
                if name.endswith("101"):
                    D = prm["eps_1"] / J * E
                elif name.endswith("105"):
                    D = np.linalg.solve(J / prm["eps_1"] * np.linalg.inv(b) + J / prm["eps_2"] * np.eye(ndim), E)
                    D_it = m.ElectricDisplacementx(st, E.reshape(ndim, 1), 0, g).ravel()
                    assert np.allclose(D, D_it, rtol=1e-6), (D, D_it)
                else:
                    D = prm["eps_2"] * E
                Ds.append(D)
                Ss.append(m.CauchyStress(st, D.reshape(ndim, 1), 0, g))
                Hs.append(m.Hessian(st, D.reshape(ndim, 1), 0, g))
            out["mat_%s_%dd_D" % (name, ndim)] = np.array(Ds)
            out["mat_%s_%dd_stress" % (name, ndim)] = np.array(Ss)
            out["mat_%s_%dd_hessian" % (name, ndim)] = np.array(Hs)


def perturbed(mesh, rng, amp=0.02):
    h = (mesh.points.max(0) - mesh.points.min(0)).min() / max(2, round(mesh.points.shape[0] ** (1.0 / mesh.points.shape[1])) - 1)
    return mesh.points + amp * h * rng.uniform(-1, 1, mesh.points.shape)


def _asm_case(out, rng, etype, p, n, matname):
    """One (mesh, material) case through the reference's AssembleForm; returns the fixture key."""
    mesh = make_mesh(etype, p, n)
    ndim = mesh.points.shape[1]
    material, prm = material_of(matname, ndim)
    material.has_low_level_dispatcher = False
    electro = matname in ELEC
    form = DisplacementPotentialFormulation(mesh) if electro else DisplacementFormulation(mesh)
    nature = "linear" if matname == "LinearElastic" else "nonlinear"
    fem_solver = FEMSolver(analysis_nature=nature, optimise=False, recompute_sparsity_pattern=True)
    Eulerx = perturbed(mesh, rng)
    Eulerp = None
    if electro:
        Eulerp = 9.0e3 * mesh.points[:, -1] / mesh.points[:, -1].max() + 10.0 * rng.uniform(-1, 1, mesh.points.shape[0])
    K, T = AssembleForm(form, mesh, material, fem_solver, Eulerx=Eulerx.copy(), Eulerp=None if Eulerp is None else Eulerp.copy())
    K = K.tocsr()
    K.sum_duplicates()
    K.sort_indices()
    tag = ("n%d" % n) if np.isscalar(n) else ("n%dx%dx%d" % tuple(n))
    key = "asm_%s%d_%s_%s" % (etype, p, tag, matname)
    fs = form.function_spaces[0]
    out[key + "_points"] = mesh.points
    out[key + "_elements"] = mesh.elements.astype(np.int64)
    out[key + "_Eulerx"] = Eulerx
    if Eulerp is not None:
        out[key + "_Eulerp"] = Eulerp
    out[key + "_Jm"] = fs.Jm
    out[key + "_AllGauss"] = fs.AllGauss
    out[key + "_Bases"] = fs.Bases
    out[key + "_K_data"] = K.data
    out[key + "_K_indices"] = K.indices
    out[key + "_K_indptr"] = K.indptr
    out[key + "_T"] = T.ravel()
    out[key + "_update"] = np.array(int(fem_solver.requires_geometry_update))
    out[key + "_prm"] = np.array([prm.get(k, 0.0) for k in ("mu", "mu1", "mu2", "mu3", "mue", "lamb", "eps_1", "eps_2", "eps_3", "eps_e")])
    # the reference's own native sparsity pattern + slot maps
    idx, iptr, dl, dg = ComputeSparsityPattern(mesh, form.nvar)
    out[key + "_sp_indices"] = idx
    out[key + "_sp_indptr"] = iptr
    if mesh.nelem * (form.nvar * mesh.elements.shape[1]) ** 2 < 400000:
        out[key + "_sp_dl"] = dl
        out[key + "_sp_dg"] = dg
    print(key, K.shape, K.nnz, "update", fem_solver.requires_geometry_update)
    return key


def gen_assembly(out):
    """Global K, T from the reference's Python path on small meshes."""
    rng = np.random.default_rng(11)
    cases = [
        ("hex", 1, 2, "NeoHookean"), ("hex", 2, 2, "NeoHookean"), ("hex", 2, 2, "MooneyRivlin"),
        ("hex", 2, 1, "NearlyIncompressibleMooneyRivlin"), ("tet", 2, 2, "LinearElastic"), ("tet", 2, 1, "NeoHookean"),
        ("hex", 3, 1, "MooneyRivlin"), ("quad", 2, 3, "NeoHookean"), ("quad", 1, 3, "MooneyRivlin"), ("tri", 2, 3, "LinearElastic"),
        ("tri", 2, 2, "NearlyIncompressibleMooneyRivlin"),
        ("hex", 2, 1, "IsotropicElectroMechanics_108"), ("hex", 1, 2, "IsotropicElectroMechanics_101"),
        ("tet", 2, 1, "IsotropicElectroMechanics_105"), ("quad", 2, 2, "IsotropicElectroMechanics_108"),
        ("tri", 2, 2, "IsotropicElectroMechanics_101"), ("hex", 3, 1, "IsotropicElectroMechanics_108"),
    ]
    names = [_asm_case(out, rng, *case) for case in cases]
    out["asm_cases"] = np.array(names)


def gen_assembly_hi(out):
    """Multi-element high-order / electro-mechanical cases: every shared node of these meshes is visited
    by 2-8 elements, so the plane-major K_e scratch + wide CSR reduction of the hex64 DMMA path and the electro hex27 path are
    compared with the reference's own K on more than one visit per node.  Kept in a separate file so that golden_assembly.npz
    (and the random stream its cases consumed) stays as committed."""
    rng = np.random.default_rng(23)
    cases = [("hex", 3, (2, 2, 1), "IsotropicElectroMechanics_108"), ("hex", 3, (2, 1, 2), "MooneyRivlin"),
             ("hex", 2, (2, 2, 2), "IsotropicElectroMechanics_108"), ("hex", 2, (2, 2, 2), "IsotropicElectroMechanics_105")]
    out["asm_cases"] = np.array([_asm_case(out, rng, *case) for case in cases])


def gen_assembly_tet3(out):
    """Cubic tetrahedra (tet20, 14 Gauss points): the p >= 3 simplex shape of BASELINE north_star ("high-order (p>=3) hex/tet").
    Small meshes (6 and 12 elements) of the reference's own tet20 node arrangement and quadrature, mechanics and electro-mechanics."""
    rng = np.random.default_rng(31)
    cases = [("tet", 3, 1, "NeoHookean"), ("tet", 3, (2, 1, 1), "MooneyRivlin"), ("tet", 3, 1, "IsotropicElectroMechanics_108"),
             ("tet", 3, 1, "LinearElastic")]
    out["asm_cases"] = np.array([_asm_case(out, rng, *case) for case in cases])


def gen_mass(out):
    """Consistent mass matrix from the reference's Python path (Assemble with analysis_type="dynamic", optimise=False):
    M = sum_e sum_g rho N N^T w |det J_X| on every displacement dof (DisplacementFormulation.GetLocalMass), zero rows for the
    potential dof of the electro-mechanical formulation.  Pins the consistent branch of __TotalConstantMassIntegrand__, which the
    explicit fixtures (lumped M only) do not reach."""
    from Florence.FiniteElements.Assembly import Assemble
    names = []
    for etype, p, n, matname in (("hex", 2, 2, "NeoHookean"), ("tet", 2, 2, "NeoHookean"), ("quad", 2, 3, "NeoHookean"),
                                 ("hex", 1, 2, "IsotropicElectroMechanics_101"), ("tet", 3, 1, "NeoHookean")):
        mesh = make_mesh(etype, p, n)
        ndim = mesh.points.shape[1]
        material, prm = material_of(matname, ndim)
        material.has_low_level_dispatcher = False
        electro = matname in ELEC
        form = DisplacementPotentialFormulation(mesh) if electro else DisplacementFormulation(mesh)
        fem_solver = FEMSolver(analysis_type="dynamic", analysis_subtype="implicit", mass_type="consistent", optimise=False,
                               analysis_nature="nonlinear")
        fem_solver.is_mass_computed = False
        fs = form.function_spaces[0]
        K, T, F, M = Assemble(fem_solver, fs, form, mesh, material, mesh.points.copy(), np.zeros(mesh.points.shape[0]))
        M = M.tocsr(); M.sum_duplicates(); M.sort_indices()
        key = "mass_%s%d_n%d_%s" % (etype, p, n, matname)
        out[key + "_points"] = mesh.points
        out[key + "_elements"] = mesh.elements.astype(np.int64)
        out[key + "_Jm"] = fs.Jm
        out[key + "_AllGauss"] = fs.AllGauss
        out[key + "_Bases"] = fs.Bases
        out[key + "_rho"] = np.array(float(material.rho))
        out[key + "_nvar"] = np.array(int(form.nvar))
        out[key + "_M_data"] = M.data
        out[key + "_M_indices"] = M.indices
        out[key + "_M_indptr"] = M.indptr
        names.append(key)
        print(key, M.shape, M.nnz, "sum", M.sum())
    out["mass_cases"] = np.array(names)


def gen_laplacian(out):
    names = []
    for etype, p, n in (("hex", 2, 2), ("hex", 4, 1), ("tet", 2, 2), ("quad", 2, 3), ("tri", 1, 3)):
        mesh = make_mesh(etype, p, n)
        ndim = mesh.points.shape[1]
        material = F.IdealDielectric(ndim, eps_1=2.35)
        material.has_low_level_dispatcher = False
        material.e = material.eps_1 * np.eye(ndim)   # what LaplacianSolver.py:41-43 sets before assembling
        form = LaplacianFormulation(mesh)
        fem_solver = FEMSolver(optimise=False, recompute_sparsity_pattern=True)
        K, T = AssembleForm(form, mesh, material, fem_solver)
        K = K.tocsr(); K.sum_duplicates(); K.sort_indices()
        key = "lap_%s%d_n%d" % (etype, p, n)
        names.append(key)
        fs = form.function_spaces[0]
        out[key + "_points"] = mesh.points
        out[key + "_elements"] = mesh.elements.astype(np.int64)
        out[key + "_Jm"] = fs.Jm
        out[key + "_AllGauss"] = fs.AllGauss
        out[key + "_e"] = np.asarray(material.e, dtype=np.float64)
        out[key + "_K_data"] = K.data
        out[key + "_K_indices"] = K.indices
        out[key + "_K_indptr"] = K.indptr
        print(key, K.shape, K.nnz)
    out["lap_cases"] = np.array(names)


def gen_explicit(out):
    """Short explicit central-difference run through the reference's own integrator (optimise=False)."""
    mesh = make_mesh("hex", 2, 2)
    ndim = 3
    material, prm = material_of("NeoHookean", ndim)
    material.has_low_level_dispatcher = False
    nsteps = 12

    def dirichlet(mesh, time_step):
        bd = np.zeros((mesh.points.shape[0], 3, time_step)) + np.nan
        bd[np.isclose(mesh.points[:, 2], 0), :, :] = 0.
        return bd

    def neumann(mesh, time_step):
        flags = np.zeros((mesh.faces.shape[0], time_step), dtype=np.uint8)
        data = np.zeros((mesh.faces.shape[0], 3, time_step))
        for i in range(mesh.faces.shape[0]):
            avg = mesh.points[mesh.faces[i, :], :].mean(0)
            if np.isclose(avg[2], mesh.points[:, 2].max()):
                data[i, 2, :] = np.linspace(0, -1e4, time_step)
                flags[i, :] = True
        return flags, data

    bc = BoundaryCondition()
    bc.SetDirichletCriteria(dirichlet, mesh, nsteps)
    bc.SetNeumannCriteria(neumann, mesh, nsteps)
    form = DisplacementFormulation(mesh)
    # dt below the CFL limit of this mesh
    h = 0.8 / 4
    dt = 0.1 * h / np.sqrt((prm["lamb"] + 2 * prm["mu"]) / 1100.0)
    fem_solver = FEMSolver(total_time=dt * nsteps, number_of_load_increments=nsteps, analysis_type="dynamic",
                           analysis_subtype="explicit", mass_type="lumped", optimise=False, print_incremental_log=False,
                           report_log_level=0)
    # record what FEMSolver.Solve hands to the integrator (lumped M, Neumann forces, first T)
    from Florence.TimeIntegrators import ExplicitStructuralDynamicIntegrator as ESDI
    rec = {}
    orig = ESDI.Solver

    def recording_solver(self, function_spaces, formulation, solver, TractionForces, M, NeumannForces, *a, **k):
        rec["T0"] = np.array(TractionForces).ravel().copy()
        rec["M"] = np.array(M).ravel().copy()
        rec["NeumannForces"] = np.array(NeumannForces).copy()
        return orig(self, function_spaces, formulation, solver, TractionForces, M, NeumannForces, *a, **k)
    ESDI.Solver = recording_solver
    sol = fem_solver.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc)
    ESDI.Solver = orig
    fs = form.function_spaces[1]
    out["exp_points"] = mesh.points
    out["exp_elements"] = mesh.elements.astype(np.int64)
    out["exp_Jm"] = fs.Jm
    out["exp_AllGauss"] = fs.AllGauss
    out["exp_Bases"] = fs.Bases
    out["exp_prm"] = np.array([prm.get(k, 0.0) for k in ("mu", "mu1", "mu2", "mu3", "mue", "lamb", "eps_1", "eps_2", "eps_3", "eps_e")])
    out["exp_rho"] = np.array(1100.0)
    out["exp_dt"] = np.array(dt)
    out["exp_nsteps"] = np.array(nsteps)
    out["exp_TotalDisp"] = sol.sol
    out["exp_columns_out"] = bc.columns_out
    out["exp_applied_dirichlet"] = np.asarray(bc.applied_dirichlet)
    out["exp_neumann"] = rec["NeumannForces"]
    out["exp_M_lumped"] = rec["M"]
    out["exp_T0"] = rec["T0"]
    print("explicit", sol.sol.shape, np.abs(sol.sol).max())


def gen_explicit_rules(out):
    """The save / termination rules of the reference's explicit integrator (ExplicitStructuralDynamicIntegrator.py:165-251), pinned by
    running it: (a) save_frequency = 3, (b) a time step above the stability limit (the growth test `abs(U.max()/(U0.max()+1e-14)) > 10`
    fires and TotalDisp is cut), (c) break_at_increment.  Same mesh, material and loading as gen_explicit; constant-vector ramp loading
    (1-D applied_dirichlet, single-column NeumannForces) so that the device-resident chunked loop of the drop-in is what gets compared."""
    ndim = 3
    nsteps_all = {"sf3": 18, "blow": 40, "brk": 15}
    for tag, nsteps in nsteps_all.items():
        mesh = make_mesh("hex", 2, 2)
        material, prm = material_of("NeoHookean", ndim)
        material.has_low_level_dispatcher = False

        def dirichlet(mesh):
            bd = np.zeros((mesh.points.shape[0], 3)) + np.nan
            bd[np.isclose(mesh.points[:, 2], 0), :] = 0.
            bd[np.isclose(mesh.points[:, 2], mesh.points[:, 2].max()), 2] = 0.01
            return bd

        def neumann(mesh):
            flags = np.zeros(mesh.faces.shape[0], dtype=np.uint8)
            data = np.zeros((mesh.faces.shape[0], 3))
            for i in range(mesh.faces.shape[0]):
                avg = mesh.points[mesh.faces[i, :], :].mean(0)
                if np.isclose(avg[0], mesh.points[:, 0].max()):
                    data[i, 0] = 2.0e3
                    flags[i] = True
            return flags, data

        bc = BoundaryCondition()
        bc.SetDirichletCriteria(dirichlet, mesh)
        bc.SetNeumannCriteria(neumann, mesh)
        form = DisplacementFormulation(mesh)
        h = 0.8 / 4
        cfl = h / np.sqrt((prm["lamb"] + 2 * prm["mu"]) / 1100.0)
        dt = (0.1 if tag != "blow" else 1.5) * cfl
        kw = dict(total_time=dt * nsteps, number_of_load_increments=nsteps, analysis_type="dynamic", analysis_subtype="explicit",
                  mass_type="lumped", optimise=False, print_incremental_log=False, report_log_level=0)
        if tag == "sf3":
            kw["memory_store_frequency"] = 3
        if tag == "brk":
            kw["break_at_increment"] = 9
        fem_solver = FEMSolver(**kw)
        from Florence.TimeIntegrators import ExplicitStructuralDynamicIntegrator as ESDI
        rec = {}
        orig = ESDI.Solver

        def recording_solver(self, function_spaces, formulation, solver, TractionForces, M, NeumannForces, NodalForces, Residual,
                             mesh_, TotalDisp, *a, **k):
            rec["T0"] = np.array(TractionForces).ravel().copy()
            rec["M"] = np.array(M).ravel().copy()
            rec["NeumannForces"] = np.array(NeumannForces).copy()
            rec["TotalDisp_shape"] = np.array(TotalDisp.shape)
            return orig(self, function_spaces, formulation, solver, TractionForces, M, NeumannForces, NodalForces, Residual, mesh_,
                        TotalDisp, *a, **k)
        ESDI.Solver = recording_solver
        try:
            sol = fem_solver.Solve(formulation=form, mesh=mesh, material=material, boundary_condition=bc)
        finally:
            ESDI.Solver = orig
        fs = form.function_spaces[1]
        pre = "rule_%s_" % tag
        if tag == "sf3":
            out["rule_points"] = mesh.points
            out["rule_elements"] = mesh.elements.astype(np.int64)
            out["rule_faces"] = mesh.faces.astype(np.int64)
            out["rule_Jm"] = fs.Jm
            out["rule_AllGauss"] = fs.AllGauss
            out["rule_Bases"] = fs.Bases
            out["rule_prm"] = np.array([prm.get(k, 0.0) for k in ("mu", "mu1", "mu2", "mu3", "mue", "lamb", "eps_1", "eps_2", "eps_3", "eps_e")])
            out["rule_columns_out"] = bc.columns_out
            out["rule_applied_dirichlet"] = np.asarray(bc.applied_dirichlet)
        out[pre + "nsteps"] = np.array(nsteps)
        out[pre + "dt"] = np.array(dt)
        out[pre + "TotalDisp"] = sol.sol
        out[pre + "TotalDisp_shape_in"] = rec["TotalDisp_shape"]
        out[pre + "number_of_load_increments"] = np.array(fem_solver.number_of_load_increments)
        out[pre + "neumann"] = rec["NeumannForces"]
        out[pre + "M"] = rec["M"]
        out[pre + "T0"] = rec["T0"]
        print(tag, "TotalDisp", sol.sol.shape, "allocated", rec["TotalDisp_shape"], "increments", fem_solver.number_of_load_increments,
              "applied_dirichlet", np.asarray(bc.applied_dirichlet).shape, "neumann", rec["NeumannForces"].shape, np.abs(sol.sol).max())


if __name__ == "__main__":
    which = sys.argv[1:] or ["tables", "materials", "assembly", "laplacian", "explicit"]
    gens = dict(tables=gen_tables, materials=gen_materials, assembly=gen_assembly, assembly_hi=gen_assembly_hi, assembly_tet3=gen_assembly_tet3, mass=gen_mass,
                laplacian=gen_laplacian,
                explicit=gen_explicit, explicit_rules=gen_explicit_rules)
    for w in which:
        out = {}
        gens[w](out)
        path = os.path.join(HERE, "golden_%s.npz" % w)
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")
