"""ctypes front end of the CPU oracle (oracle/florence_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  florence_b200/ never does.

The functions mirror the reference's low-level entry points
(Florence/FiniteElements/Assembly/_Assembly_/*.pyx) on plain numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

MATERIAL_NUMBERS = {
    # Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyExplicit_DF_DPF_.pyx:72-109
    "ExplicitMooneyRivlin": 0,
    "NeoHookean": 1,
    "MooneyRivlin": 2,
    "NearlyIncompressibleMooneyRivlin": 3,
    "IsotropicElectroMechanics_101": 4,
    "IsotropicElectroMechanics_105": 5,
    "IsotropicElectroMechanics_108": 8,
    "ExplicitIsotropicElectroMechanics_108": 9,
    "LinearElastic": 10,
    "IncrementalLinearElastic": 10,
}
PARAM_ORDER = ("mu", "mu1", "mu2", "mu3", "mue", "lamb", "eps_1", "eps_2", "eps_3", "eps_e")


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not (os.path.exists(os.path.join(_BUILD, "liboracle.so")) and os.path.exists(os.path.join(_BUILD, "liboracle_fast.so"))):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)


_libs = {}


def _lib(fast=False):
    key = "fast" if fast else "strict"
    if key not in _libs:
        path = os.path.join(_BUILD, "liboracle_fast.so" if fast else "liboracle.so")
        if not os.path.exists(path):
            build()
        _libs[key] = C.CDLL(path)
    return _libs[key]


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def params(**kw):
    """Pack material constants in the reference C-signature order (mu, mu1, mu2, mu3, mue, lamb, eps_1, eps_2, eps_3, eps_e)."""
    p = np.zeros(10)
    for k, v in kw.items():
        p[PARAM_ORDER.index(k)] = v
    return p


def hessian_size(material_number, ndim):
    electro = material_number in (4, 5, 8, 9)
    return (6 if ndim == 3 else 3) + (ndim if electro else 0)


def material_point(material_number, F, E=None, prm=None, want_hessian=True):
    """sigma (d,d), H (hs,hs) [, D (d)] at one point -- LLDispatch/CythonSource/_<Material>_.h::_KineticMeasures_."""
    F = _f64(F)
    d = F.shape[0]
    E = np.zeros(3) if E is None else _f64(np.concatenate([np.ravel(E), np.zeros(3 - d)]))
    hs = hessian_size(material_number, d)
    D = np.zeros(3)
    S = np.zeros((d, d))
    H = np.zeros((hs, hs)) if want_hessian else None
    rc = _lib().flo_material_point(C.c_int(material_number), C.c_int(d), _p(F), _p(E), _p(_f64(prm)), _p(D), _p(S), _p(H))
    if rc:
        raise NotImplementedError("material %d not restated" % material_number)
    return D[:d], S, H


def assemble_explicit(points, elements, Eulerx, Eulerp, Jm, AllGauss, nvar, prm, material_number, formulation_number=0,
                      elem_range=None, T=None, fast=False):
    """_LowLevelAssemblyExplicit_DF_DPF_ (.pyx:44-152): returns T (nnode*nvar)."""
    points, Eulerx, Jm = _f64(points), _f64(Eulerx), _f64(Jm)
    AllGauss = _f64(AllGauss).ravel()
    elements = np.ascontiguousarray(elements, dtype=np.uint64)
    nnode, ndim = points.shape
    nelem, npe = elements.shape
    Eulerp = np.zeros(nnode) if Eulerp is None else _f64(Eulerp).ravel()
    if T is None:
        T = np.zeros(nnode * nvar)
    e0, e1 = (0, nelem) if elem_range is None else elem_range
    rc = _lib(fast).flo_assemble_explicit(_p(points), _p(elements), _p(Eulerx), _p(Eulerp), _p(Jm), _p(AllGauss), C.c_int64(ndim),
                                          C.c_int64(nvar), C.c_int64(AllGauss.shape[0]), C.c_int64(e0), C.c_int64(e1), C.c_int64(npe),
                                          _p(T), _p(_f64(prm)), C.c_int(material_number), C.c_int(formulation_number))
    if rc:
        raise NotImplementedError("material %d not restated" % material_number)
    return T


def assemble_implicit(points, elements, Eulerx, Eulerp, Jm, AllGauss, nvar, H_VoigtSize, requires_geometry_update, prm,
                      material_number, mode="coo", pattern=None, elem_range=None, out=None, fast=False):
    """_LowLevelAssemblyDF_/_DPF_ (.pyx:53-160).

    mode "coo": returns (I, J, V, T) with ndof^2*nelem triplets (recompute_sparsity_pattern=True);
    mode "csr": pattern=(indices, indptr, data_local_indices, data_global_indices) -> (V, T);
    mode "csr_search": pattern=(indices, indptr, sorted_elements, sorter) -> (V, T) (squeeze_sparsity_pattern=True).
    """
    points, Eulerx, Jm = _f64(points), _f64(Eulerx), _f64(Jm)
    AllGauss = _f64(AllGauss).ravel()
    elements = np.ascontiguousarray(elements, dtype=np.uint64)
    nnode, ndim = points.shape
    nelem, npe = elements.shape
    ndof = nvar * npe
    Eulerp = np.zeros(nnode) if Eulerp is None else _f64(Eulerp).ravel()
    e0, e1 = (0, nelem) if elem_range is None else elem_range
    dl = dg = se = so = None
    if mode == "coo":
        imode = 0
        if out is None:
            I = np.zeros(ndof * ndof * nelem, np.int32)
            J = np.zeros(ndof * ndof * nelem, np.int32)
            V = np.zeros(ndof * ndof * nelem)
            T = np.zeros(nnode * nvar)
        else:
            I, J, V, T = out
    else:
        indices, indptr = pattern[0], pattern[1]
        I, J = np.ascontiguousarray(indptr, np.int32), np.ascontiguousarray(indices, np.int32)
        if out is None:
            V = np.zeros(J.shape[0])
            T = np.zeros(nnode * nvar)
        else:
            V, T = out
        if mode == "csr":
            imode = 1
            dl, dg = np.ascontiguousarray(pattern[2], np.int32), np.ascontiguousarray(pattern[3], np.int32)
        else:
            imode = 2
            se, so = np.ascontiguousarray(pattern[2], np.uint64), np.ascontiguousarray(pattern[3], np.int64)
    rc = _lib(fast).flo_assemble_implicit(_p(points), _p(elements), _p(Eulerx), _p(Eulerp), _p(Jm), _p(AllGauss), C.c_int64(ndim),
                                          C.c_int64(nvar), C.c_int64(AllGauss.shape[0]), C.c_int64(e0), C.c_int64(e1), C.c_int64(npe),
                                          C.c_int64(H_VoigtSize), C.c_int64(int(requires_geometry_update)), _p(I), _p(J), _p(V), _p(T),
                                          C.c_int(imode), _p(dl), _p(dg), _p(se), _p(so), _p(_f64(prm)), C.c_int(material_number))
    if rc:
        raise NotImplementedError("material %d not restated" % material_number)
    return (I, J, V, T) if mode == "coo" else (V, T)


def element_implicit(X, x, phi, Jm, AllGauss, nvar, H, update, geometric, prm, material_number):
    """K_e (ndof,ndof), t_e (ndof) of one element (body of _LowLevelAssemblyDF_.h:69-131)."""
    X, x, Jm = _f64(X), _f64(x), _f64(Jm)
    AllGauss = _f64(AllGauss).ravel()
    npe, ndim = X.shape
    phi = np.zeros(npe) if phi is None else _f64(phi)
    ndof = nvar * npe
    K = np.zeros((ndof, ndof))
    Tr = np.zeros(ndof)
    rc = _lib().flo_element_implicit(_p(X), _p(x), _p(phi), _p(Jm), _p(AllGauss), C.c_int(ndim), C.c_int(nvar), C.c_int(AllGauss.shape[0]),
                                     C.c_int(npe), C.c_int(H), C.c_int(int(update)), C.c_int(int(geometric)), _p(_f64(prm)),
                                     C.c_int(material_number), _p(K), _p(Tr))
    if rc:
        raise NotImplementedError("material %d not restated" % material_number)
    return K, Tr


def assemble_laplacian(points, elements, Jm, AllGauss, e_tensor, is_hessian_symmetric=True, mode="coo", pattern=None,
                       elem_range=None, fast=False):
    """_LowLevelAssemblyPerfectLaplacian_ (.pyx); e_tensor is what the wrapper passes, i.e. -material.e (.pyx:81)."""
    points, Jm = _f64(points), _f64(Jm)
    AllGauss = _f64(AllGauss).ravel()
    elements = np.ascontiguousarray(elements, dtype=np.uint64)
    nnode, ndim = points.shape
    nelem, npe = elements.shape
    e0, e1 = (0, nelem) if elem_range is None else elem_range
    dl = dg = se = so = None
    if mode == "coo":
        imode = 0
        I = np.zeros(npe * npe * nelem, np.int32)
        J = np.zeros(npe * npe * nelem, np.int32)
        V = np.zeros(npe * npe * nelem)
    else:
        I, J = np.ascontiguousarray(pattern[1], np.int32), np.ascontiguousarray(pattern[0], np.int32)
        V = np.zeros(J.shape[0])
        if mode == "csr":
            imode = 1
            dl, dg = np.ascontiguousarray(pattern[2], np.int32), np.ascontiguousarray(pattern[3], np.int32)
        else:
            imode = 2
            se, so = np.ascontiguousarray(pattern[2], np.uint64), np.ascontiguousarray(pattern[3], np.int64)
    _lib(fast).flo_assemble_laplacian(_p(points), _p(elements), _p(Jm), _p(AllGauss), C.c_int64(ndim), C.c_int64(AllGauss.shape[0]),
                                      C.c_int64(e0), C.c_int64(e1), C.c_int64(npe), _p(I), _p(J), _p(V), _p(_f64(e_tensor)),
                                      C.c_int(int(is_hessian_symmetric)), C.c_int(imode), _p(dl), _p(dg), _p(se), _p(so))
    return (I, J, V) if mode == "coo" else V


def assemble_mass(points, elements, Bases, Jm, AllGauss, nvar, rho, mass_type="lumped"):
    """__TotalConstantMassIntegrand__ (_MassIntegrand_.pyx:192-349), generic integrator (.h:249-395)."""
    points, Jm, Bases = _f64(points), _f64(Jm), _f64(Bases)
    AllGauss = _f64(AllGauss).ravel()
    elements = np.ascontiguousarray(elements, dtype=np.uint64)
    nnode, ndim = points.shape
    nelem, npe = elements.shape
    ndof = nvar * npe
    if mass_type == "lumped":
        mass = np.zeros(nnode * nvar)
        I = J = V = None
    else:
        mass = None
        I = np.zeros(ndof * ndof * nelem, np.int32)
        J = np.zeros(ndof * ndof * nelem, np.int32)
        V = np.zeros(ndof * ndof * nelem)
    _lib().flo_assemble_mass(_p(points), _p(elements), _p(Bases), _p(Jm), _p(AllGauss), C.c_int64(ndim), C.c_int64(nvar),
                             C.c_int64(AllGauss.shape[0]), C.c_int64(0), C.c_int64(nelem), C.c_int64(npe), C.c_double(rho),
                             C.c_int(0 if mass_type == "lumped" else 1), _p(mass), _p(I), _p(J), _p(V))
    return mass if mass_type == "lumped" else (I, J, V)


def sparsity_pattern(elements, nnode, nvar, with_data_indices=True):
    """ComputeSparsityPattern (.pyx:44-112): (indices, indptr[, data_local_indices, data_global_indices])."""
    elements = np.ascontiguousarray(elements, dtype=np.uint64)
    nelem, npe = elements.shape
    lib = _lib()
    lib.flo_sparsity_pattern.restype = C.c_int64
    indptr = np.zeros(nnode * nvar + 1, np.int32)
    nnz = lib.flo_sparsity_pattern(_p(elements), C.c_int64(nelem), C.c_int64(npe), C.c_int64(nnode), C.c_int64(nvar), _p(indptr), None)
    indices = np.zeros(nnz, np.int32)
    lib.flo_sparsity_pattern(_p(elements), C.c_int64(nelem), C.c_int64(npe), C.c_int64(nnode), C.c_int64(nvar), _p(indptr), _p(indices))
    if not with_data_indices:
        return indices, indptr
    cap = (nvar * npe) ** 2
    dl = np.zeros(cap * nelem, np.int32)
    dg = np.zeros(cap * nelem, np.int32)
    lib.flo_data_indices(_p(elements), C.c_int64(nelem), C.c_int64(npe), C.c_int64(nvar), _p(indptr), _p(indices), _p(dl), _p(dg), None, None)
    return indices, indptr, dl, dg


def element_sorter(elements):
    """(sorted_elements uint64, sorter int64): mesh.sorted_elements / mesh.element_sorter used in squeeze mode."""
    elements = np.ascontiguousarray(elements, dtype=np.uint64)
    nelem, npe = elements.shape
    so = np.zeros((nelem, npe), np.int64)
    se = np.zeros((nelem, npe), np.uint64)
    _lib().flo_data_indices(_p(elements), C.c_int64(nelem), C.c_int64(npe), C.c_int64(1), None, None, None, None, _p(so), _p(se))
    return se, so


def explicit_central_difference(T_of, X, M, F_ext_of, dt, nsteps, fixed_mask, applied_dirichlet_of=None, save_every=1, contact_of=None):
    """Lumped-mass central-difference loop, mechanics only: numpy restatement of
    Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:57-188 (zero initial U0, V0).

    T_of(Eulerx) -> internal force (nnode*ndim); F_ext_of(inc) -> nodal forces; fixed_mask bool (nnode*ndim).
    contact_of(Eulerx) -> contact tractions added after every in-loop assembly (:190-197; not to the initial T).
    Returns (snapshots list of U arrays, final Eulerx, final T).
    """
    nnode, ndim = X.shape
    ndof = nnode * ndim
    invM = np.reciprocal(M)
    free = ~fixed_mask
    T = T_of(X.copy())
    U0 = np.zeros(ndof)
    V0 = np.zeros(ndof)
    A0 = (F_ext_of(0) - T) * invM                        # :63-66
    U00 = U0 - dt * V0 + (dt ** 2 / 2.) * A0             # :79
    U0 = np.where(free, U0, 0.0)                         # :80-81 UpdateFreeMechanicalDoFs zeroes the fixed dofs
    U00 = np.where(free, U00, 0.0)
    snaps = []
    Eulerx = X.copy()
    for inc in range(2, nsteps):
        R = F_ext_of(inc) - T                            # :131-132
        R = R + (2. / dt ** 2) * M * U0 - (1. / dt ** 2) * M * U00   # :135
        U = dt ** 2 * invM * R                           # :136
        U = np.where(free, U, 0.0)                       # :137
        inc_dir = np.zeros(ndof)
        if applied_dirichlet_of is not None:
            inc_dir[fixed_mask] = applied_dirichlet_of(inc)          # :138-139
        Eulerx = X + (U + inc_dir).reshape(nnode, ndim)  # :156-157
        if inc % save_every == 0:
            snaps.append((U + inc_dir).copy())
        U00, U0 = U0, U                                  # :183-184
        T = T_of(Eulerx)                                 # :188
        if contact_of is not None:
            T = T + contact_of(Eulerx)                   # :190-197
    return snaps, Eulerx, T


# ------------------------------------------------------------------------------------------------ the reference's own natives
_REF_PATH = os.path.join(_HERE, "_ref", "libflorence_ref.so")


def build_ref(reference_root="/root/reference"):
    """Compile the reference's Fastor-free native headers from where they lie (oracle/Makefile target `ref`).
    Only possible where the reference checkout exists (the build container); the .so then travels with the repo."""
    if os.path.isdir(reference_root):
        subprocess.check_call(["make", "-C", _HERE, "ref", "REF=" + reference_root], stdout=subprocess.DEVNULL)
    return os.path.exists(_REF_PATH)


def ref_available():
    return os.path.exists(_REF_PATH)


def ref_sparsity_pattern(elements, nnode, nvar):
    """ComputeSparsityPattern executed by the REFERENCE's native code (_ComputeSparsityPattern_, _ComputeDataIndices_); the numpy
    glue around the two calls follows ComputeSparsityPattern.pyx:44-112 line by line."""
    lib = C.CDLL(_REF_PATH)
    els = np.asarray(elements)
    nelem, nodeperelem = els.shape
    to_c_elements = np.copy(els.astype(np.int32))
    sorter = np.ascontiguousarray(np.argsort(els, axis=1).astype(np.int64))
    to_c_elements = np.ascontiguousarray(to_c_elements[np.arange(nelem)[:, None], sorter])
    flat = np.ascontiguousarray(to_c_elements.ravel())
    idx_sort = np.argsort(flat, kind="stable").astype(np.int32)
    sorted_elements = flat[idx_sort]
    elem_container = np.ascontiguousarray((idx_sort // nodeperelem).astype(np.int32))
    idx_start = np.zeros(nnode + 1, dtype=np.int32)
    idx_start[:-1] = np.unique(sorted_elements, return_index=True)[1].astype(np.int32)
    idx_start[-1] = elem_container.shape[0]
    counts = np.zeros(idx_start.shape[0] - 1, dtype=np.int32)
    indices = np.zeros(int((nvar * nodeperelem) ** 2 * nelem), dtype=np.int32)
    lib.ref_compute_sparsity_pattern.restype = C.c_int
    nnz = lib.ref_compute_sparsity_pattern(_p(flat), _p(idx_start), _p(elem_container), C.c_int(nvar), C.c_int(nnode), C.c_int(nelem),
                                           C.c_int(nodeperelem), C.c_int(idx_start.shape[0]), _p(counts), _p(indices))
    counts = np.repeat(counts, nvar)
    all_ndof = nnode * nvar
    indptr = np.zeros(nnode * nvar + 1, dtype=np.int32)
    indptr[1:] = np.cumsum(np.minimum(counts.astype(np.int64) * nvar, all_ndof)).astype(np.int32)
    indices = np.ascontiguousarray(indices[:nnz])
    cap = (nvar * nodeperelem) ** 2
    dl = np.zeros(cap * nelem, dtype=np.int32)
    dg = np.zeros(cap * nelem, dtype=np.int32)
    lib.ref_compute_data_indices(_p(indices), _p(indptr), C.c_int(nelem), C.c_int(nvar), C.c_int(nodeperelem), _p(to_c_elements), _p(sorter),
                                 _p(dl), _p(dg))
    return indices, indptr, dl, dg


def ref_csr_scatter(Ke_all, dl, dg, nnz):
    """SparseAssemblyNativeCSR_ (the reference's code) over all elements, in element order: V (nnz)."""
    lib = C.CDLL(_REF_PATH)
    Ke_all = _f64(Ke_all)
    nelem = Ke_all.shape[0]
    cap = Ke_all.shape[1]
    V = np.zeros(nnz)
    for e in range(nelem):
        row = np.ascontiguousarray(Ke_all[e])
        lib.ref_sparse_assembly_csr(_p(row), _p(dl), _p(dg), C.c_int(e), C.c_int(cap), _p(V))
    return V
