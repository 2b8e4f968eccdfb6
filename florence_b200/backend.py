"""Device-side assembly handle: torch tensors in, torch tensors out, all work done by libflorence_b200.so.

torch is plumbing only (device memory, streams, DLPack exchange); every number is produced by the CUDA kernels behind the C ABI.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ExplicitCtrl, Material, MeshDesc, UpdateArgs, check

# Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyExplicit_DF_DPF_.pyx:72-109
MATERIAL_NUMBERS = {
    "ExplicitMooneyRivlin": 0,
    "NeoHookean": 1,
    "MooneyRivlin": 2,
    "NearlyIncompressibleMooneyRivlin": 3,
    "IsotropicElectroMechanics_101": 4,
    "IsotropicElectroMechanics_105": 5,
    "IsotropicElectroMechanics_106": 6,
    "IsotropicElectroMechanics_107": 7,
    "IsotropicElectroMechanics_108": 8,
    "ExplicitIsotropicElectroMechanics_108": 9,
    "LinearElastic": 10,
    "IncrementalLinearElastic": 10,
}
ELECTRO_NUMBERS = (4, 5, 6, 7, 8, 9)


def to_device(x, dtype, device):
    """numpy array / torch tensor / any DLPack exporter (cupy, jax, ...) -> contiguous torch tensor on `device`."""
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        if x.dtype == np.uint64:
            # torch's uint64 support is partial; the bit pattern is what the C ABI reads
            x = x.view(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=True).contiguous()


def make_material(material_number, rho=0.0, **constants):
    m = Material()
    m.material_number = int(material_number)
    m.rho = float(rho if rho is not None else 0.0)
    for k, v in constants.items():
        setattr(m, k, float(v))
    return m


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class AssemblyHandle(object):
    """Caches a mesh + function-space tables on the device (fl_create) and exposes the assembly entry points."""

    def __init__(self, points, elements, Jm, AllGauss, Bases=None, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.FlorenceB200Error("florence_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        with torch.cuda.device(self.device):
            pts = to_device(points, torch.float64, self.device)
            els = to_device(elements, torch.int64, self.device)
            jm = to_device(Jm, torch.float64, self.device)
            gw = to_device(AllGauss, torch.float64, self.device).reshape(-1)
            bs = None if Bases is None else to_device(Bases, torch.float64, self.device)
            self.nnode, self.ndim = int(pts.shape[0]), int(pts.shape[1])
            self.nelem, self.npe = int(els.shape[0]), int(els.shape[1])
            self.ngauss = int(gw.shape[0])
            if tuple(jm.shape) != (self.ndim, self.npe, self.ngauss):
                raise ValueError("Jm must be (ndim x nodeperelem x ngauss), got %s" % (tuple(jm.shape),))
            desc = MeshDesc(self.ndim, self.npe, self.ngauss, 0, self.nelem, self.nnode, pts.data_ptr(), els.data_ptr(),
                            0 if bs is None else bs.data_ptr(), jm.data_ptr(), gw.data_ptr())
            h = C.c_void_p()
            torch.cuda.synchronize(self.device)
            check(self.lib.fl_create(C.byref(desc), C.byref(h)))
            self._h = h
        self._pattern_nvar = {}
        self.nnz = {}

    def close(self):
        if getattr(self, "_h", None):
            self.lib.fl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------------- explicit
    def assemble_explicit(self, Eulerx, Eulerp, material, formulation_number=0, out=None):
        nvar = self.ndim + (1 if formulation_number == 1 else 0)
        x = to_device(Eulerx, torch.float64, self.device)
        p = None if Eulerp is None else to_device(Eulerp, torch.float64, self.device).reshape(-1)
        T = out if out is not None else torch.empty(self.nnode * nvar, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.fl_assemble_explicit(self._h, _ptr(x), _ptr(p), C.byref(material), formulation_number, _ptr(T), _stream()))
        return T

    # ---------------------------------------------------------------------------------------------- pattern
    def build_pattern(self, nvar):
        if nvar not in self.nnz:
            nnz = C.c_int64(0)
            with torch.cuda.device(self.device):
                check(self.lib.fl_pattern_build(self._h, nvar, C.byref(nnz)))
            self.nnz[nvar] = int(nnz.value)
        return self.nnz[nvar]

    def sparsity_pattern(self, nvar, with_data_indices=False):
        """(indices, indptr[, data_local_indices, data_global_indices]) int32 device tensors, reference ordering."""
        nnz = self.build_pattern(nvar)
        if nvar not in self._pattern_nvar:
            indptr = torch.empty(self.nnode * nvar + 1, dtype=torch.int32, device=self.device)
            indices = torch.empty(nnz, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                check(self.lib.fl_pattern_export(self._h, nvar, _ptr(indptr), _ptr(indices), _stream()))
            self._pattern_nvar[nvar] = (indices, indptr)
        indices, indptr = self._pattern_nvar[nvar]
        if not with_data_indices:
            return indices, indptr
        cap = (self.npe * nvar) ** 2
        dl = torch.empty(cap * self.nelem, dtype=torch.int32, device=self.device)
        dg = torch.empty(cap * self.nelem, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.fl_pattern_export_data_indices(self._h, nvar, _ptr(dl), _ptr(dg), _stream()))
        return indices, indptr, dl, dg

    # ---------------------------------------------------------------------------------------------- penalty contact
    def set_contact(self, surface_nodes, plane_normal, distance, kappa, contact_gap_tolerance=1e-6):
        """Members of ExplicitPenaltyContactFormulation (:14-24) + the unique nodes of mesh.faces / mesh.edges (:157-160).
        surface_nodes=None switches contact off."""
        with torch.cuda.device(self.device):
            if surface_nodes is None:
                check(self.lib.fl_set_contact(self._h, None, 0, None, 0.0, 0.0, 0.0, _stream()))
                return
            ids = to_device(surface_nodes, torch.int32, self.device).reshape(-1)
            nrm = (C.c_double * 3)(*[float(v) for v in np.asarray(plane_normal, dtype=np.float64).ravel()[:self.ndim]])
            check(self.lib.fl_set_contact(self._h, _ptr(ids), ids.numel(), C.cast(nrm, C.c_void_p), float(distance), float(kappa),
                                          float(contact_gap_tolerance), _stream()))

    def assemble_contact(self, Eulerx, out=None, accumulate=False):
        """AssembleTractions (:145-184): T_contact (nnode*ndim), or `out += T_contact` when accumulate."""
        x = to_device(Eulerx, torch.float64, self.device)
        T = out if out is not None else torch.empty(self.nnode * self.ndim, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.fl_assemble_contact(self._h, _ptr(x), _ptr(T), 1 if accumulate else 0, _stream()))
        return T

    # ---------------------------------------------------------------------------------------------- Dirichlet reduction
    def dirichlet_build(self, nvar, columns_out):
        """Prescribed dofs (ascending) -> (n_in, nnz_b); BoundaryCondition.py:375-396 builds the same two lists on the host."""
        self.build_pattern(nvar)
        co = to_device(columns_out, torch.int32, self.device).reshape(-1)
        n_in, nnz_b = C.c_int64(0), C.c_int64(0)
        with torch.cuda.device(self.device):
            check(self.lib.fl_dirichlet_build(self._h, nvar, _ptr(co) if co.numel() else None, co.numel(), C.byref(n_in), C.byref(nnz_b)))
        self._dirichlet = (nvar, int(n_in.value), int(nnz_b.value), co.numel())
        return int(n_in.value), int(nnz_b.value)

    def dirichlet_pattern(self):
        """(indices_b, indptr_b, columns_in) int32 device tensors of K[columns_in,:][:,columns_in]."""
        nvar, n_in, nnz_b, _ = self._dirichlet
        indptr = torch.empty(n_in + 1, dtype=torch.int32, device=self.device)
        indices = torch.empty(nnz_b, dtype=torch.int32, device=self.device)
        cin = torch.empty(n_in, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.fl_dirichlet_export(self._h, _ptr(indptr), _ptr(indices), _ptr(cin), _stream()))
        return indices, indptr, cin

    def dirichlet_apply(self, V=None, F=None, applied=None, load_factor=1.0, want_values=True, want_reduced_force=True, out=None):
        """One pass over the free rows: V_b = values of V[in][:, in]; F[in] -= V[in][:, out(nz)] @ applied(nz) * load_factor (in
        place); F_b = F[in].  Returns (V_b or None, F_b or None)."""
        nvar, n_in, nnz_b, n_out = self._dirichlet
        Vb = Fb = None
        if out is not None:
            Vb, Fb = out
        if want_values and Vb is None:
            Vb = torch.empty(nnz_b, dtype=torch.float64, device=self.device)
        if want_reduced_force and Fb is None:
            Fb = torch.empty(n_in, dtype=torch.float64, device=self.device)
        if applied is not None:
            applied = to_device(applied, torch.float64, self.device).reshape(-1)
            if applied.numel() != n_out:
                raise ValueError("AppliedDirichlet has %d entries, columns_out has %d" % (applied.numel(), n_out))
        if F is not None and (F.dtype != torch.float64 or not F.is_contiguous() or F.numel() != self.nnode * nvar):
            raise ValueError("F must be a contiguous float64 device tensor of nvar*nnode entries")
        with torch.cuda.device(self.device):
            check(self.lib.fl_dirichlet_apply(self._h, _ptr(V) if V is not None else None, _ptr(Vb) if want_values else None,
                                              _ptr(applied) if applied is not None else None, float(load_factor),
                                              _ptr(F) if F is not None else None, _ptr(Fb) if want_reduced_force else None, _stream()))
        return (Vb if want_values else None), (Fb if want_reduced_force else None)

    # ---------------------------------------------------------------------------------------------- implicit
    def assemble_implicit(self, Eulerx, Eulerp, material, formulation_number=0, requires_geometry_update=True, mode="csr",
                          out=None, with_indices=True):
        """mode "coo": (I, J, V, T);  mode "csr": (V, T) aligned with sparsity_pattern(nvar)."""
        nvar = self.ndim + (1 if formulation_number == 1 else 0)
        ndof = self.npe * nvar
        x = to_device(Eulerx, torch.float64, self.device)
        p = None if Eulerp is None else to_device(Eulerp, torch.float64, self.device).reshape(-1)
        I = J = None
        if mode == "coo":
            n = ndof * ndof * self.nelem
            if out is not None:
                I, J, V, T = out
            else:
                if with_indices:
                    I = torch.empty(n, dtype=torch.int32, device=self.device)
                    J = torch.empty(n, dtype=torch.int32, device=self.device)
                V = torch.empty(n, dtype=torch.float64, device=self.device)
                T = torch.empty(self.nnode * nvar, dtype=torch.float64, device=self.device)
            imode = _lib.FL_MODE_COO
        else:
            nnz = self.build_pattern(nvar)
            if out is not None:
                V, T = out
            else:
                V = torch.empty(nnz, dtype=torch.float64, device=self.device)
                T = torch.empty(self.nnode * nvar, dtype=torch.float64, device=self.device)
            imode = _lib.FL_MODE_CSR
        with torch.cuda.device(self.device):
            check(self.lib.fl_assemble_implicit(self._h, _ptr(x), _ptr(p), C.byref(material), formulation_number,
                                                1 if requires_geometry_update else 0, imode, _ptr(I), _ptr(J), _ptr(V), _ptr(T), _stream()))
        return (I, J, V, T) if mode == "coo" else (V, T)

    def assemble_laplacian(self, e_tensor, is_hessian_symmetric=True, mode="csr"):
        e = np.ascontiguousarray(np.asarray(e_tensor, dtype=np.float64))
        if e.shape != (self.ndim, self.ndim):
            raise ValueError("Permittivity tensor has to have a size of (ndim x ndim)")
        I = J = None
        if mode == "coo":
            n = self.npe * self.npe * self.nelem
            I = torch.empty(n, dtype=torch.int32, device=self.device)
            J = torch.empty(n, dtype=torch.int32, device=self.device)
            V = torch.empty(n, dtype=torch.float64, device=self.device)
            imode = _lib.FL_MODE_COO
        else:
            V = torch.empty(self.build_pattern(1), dtype=torch.float64, device=self.device)
            imode = _lib.FL_MODE_CSR
        with torch.cuda.device(self.device):
            check(self.lib.fl_assemble_laplacian(self._h, e.ctypes.data_as(C.c_void_p), 1 if is_hessian_symmetric else 0, imode,
                                                 _ptr(I), _ptr(J), _ptr(V), _stream()))
            torch.cuda.current_stream().synchronize()  # e_tensor is read from host memory asynchronously
        return (I, J, V) if mode == "coo" else V

    def assemble_mass(self, rho, nvar, mass_type="lumped", mode="coo"):
        if mass_type == "lumped":
            M = torch.empty(self.nnode * nvar, dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                check(self.lib.fl_assemble_mass(self._h, float(rho), nvar, 0, 0, _ptr(M), None, None, None, _stream()))
            return M
        ndof = self.npe * nvar
        I = J = None
        if mode == "coo":
            n = ndof * ndof * self.nelem
            I = torch.empty(n, dtype=torch.int32, device=self.device)
            J = torch.empty(n, dtype=torch.int32, device=self.device)
            V = torch.empty(n, dtype=torch.float64, device=self.device)
            imode = _lib.FL_MODE_COO
        else:
            V = torch.empty(self.build_pattern(nvar), dtype=torch.float64, device=self.device)
            imode = _lib.FL_MODE_CSR
        with torch.cuda.device(self.device):
            check(self.lib.fl_assemble_mass(self._h, float(rho), nvar, 1, imode, None, _ptr(I), _ptr(J), _ptr(V), _stream()))
        return (I, J, V) if mode == "coo" else V

    # ---------------------------------------------------------------------------------------------- explicit time loop
    def explicit_steps(self, material, dt, nsteps, increment, M, fext, fixed_mask, inc_dirichlet, U0, U00, Eulerx, T,
                       fext_scale0=0.0, fext_scale_step=0.0, incd_scale0=1.0, incd_scale_step=0.0, full_status=False):
        """`nsteps` central-difference increments on the device.  Returns the status bits (0 = fine, bit 0 = NaN, bit 1 = the
        reference's growth test fired); with full_status, (bits, increment of the first detection)."""
        ctrl = ExplicitCtrl(float(dt), float(fext_scale0), float(fext_scale_step), float(incd_scale0), float(incd_scale_step),
                            int(increment), int(nsteps))
        status = (C.c_int32 * 2)(0, 0)
        with torch.cuda.device(self.device):
            check(self.lib.fl_explicit_steps(self._h, C.byref(material), C.byref(ctrl), _ptr(M), _ptr(fext), _ptr(fixed_mask),
                                             _ptr(inc_dirichlet), _ptr(U0), _ptr(U00), _ptr(Eulerx), _ptr(T), status, _stream()))
        return (int(status[0]), int(status[1])) if full_status else int(status[0])

    def explicit_update(self, dt, fext_scale, M, fext, fixed_mask, inc_dirichlet, T, U0, U00, Eulerx, status=None, growth_keys=None,
                        incd_scale=1.0, use_element_forces=False, write_T=False, iface_slot=None, T_iface=None):
        """One update (fl_explicit_update).  status: int32[2] device tensor; growth_keys: int64[2] device tensor."""
        a = UpdateArgs(float(dt), float(fext_scale), float(incd_scale), M.data_ptr(), 0 if fext is None else fext.data_ptr(),
                       0 if fixed_mask is None else fixed_mask.data_ptr(), 0 if inc_dirichlet is None else inc_dirichlet.data_ptr(),
                       T.data_ptr(), 0 if iface_slot is None else iface_slot.data_ptr(), 0 if T_iface is None else T_iface.data_ptr(),
                       U0.data_ptr(), U00.data_ptr(), Eulerx.data_ptr(), 0 if status is None else status.data_ptr(),
                       0 if growth_keys is None else growth_keys.data_ptr(), 1 if use_element_forces else 0, 1 if write_T else 0)
        with torch.cuda.device(self.device):
            check(self.lib.fl_explicit_update(self._h, C.byref(a), _stream()))

    def explicit_check(self, growth_keys, increment, status):
        with torch.cuda.device(self.device):
            check(self.lib.fl_explicit_check(self._h, _ptr(growth_keys), int(increment), _ptr(status), _stream()))

    def explicit_forces(self, Eulerx, material, e0=0, e1=None):
        """Per-element internal forces of elements [e0, e1) into the handle's scratch."""
        with torch.cuda.device(self.device):
            check(self.lib.fl_explicit_forces(self._h, _ptr(Eulerx), C.byref(material), int(e0), int(self.nelem if e1 is None else e1), _stream()))

    def gather_pack_nodes(self, nvar, node_ids, buf):
        with torch.cuda.device(self.device):
            check(self.lib.fl_gather_pack_nodes(self._h, nvar, _ptr(node_ids), node_ids.numel(), _ptr(buf), _stream()))

    def gather_nodes(self, nvar, T):
        with torch.cuda.device(self.device):
            check(self.lib.fl_gather_nodes(self._h, nvar, _ptr(T), _stream()))

    # ---------------------------------------------------------------------------------------------- owned CSR row block
    def row_block(self, nvar, V, owned_nodes, node_map):
        """(indptr_block int64 [n_owned*nvar+1], cols_global int64, vals): the CSR rows of `owned_nodes` (ascending local ids) of the
        locally assembled V with GLOBAL column dof numbers, produced on the device (fl_row_block_build / fl_row_block_emit).
        indptr_block and cols_global belong to the sparsity pattern and are built once per (nvar, owned set); every call emits the
        values -- as a zero-copy slice of V when the owned nodes are a contiguous range (slab partitions), else compacted."""
        own = to_device(owned_nodes, torch.int32, self.device).reshape(-1)
        key = ("rowblock", nvar, own.data_ptr(), own.numel())
        cache = getattr(self, "_row_block_cache", None)
        if cache is None or cache[0] != key:
            nmap = to_device(node_map, torch.int64, self.device).reshape(-1)
            indptr = torch.empty(own.numel() * nvar + 1, dtype=torch.int64, device=self.device)
            nnz = C.c_int64(0)
            with torch.cuda.device(self.device):
                check(self.lib.fl_row_block_build(self._h, nvar, _ptr(own), own.numel(), _ptr(indptr), C.byref(nnz), _stream()))
                cols = torch.empty(int(nnz.value), dtype=torch.int64, device=self.device)
                check(self.lib.fl_row_block_emit(self._h, nvar, None, _ptr(own), own.numel(), _ptr(nmap), _ptr(indptr), _ptr(cols), None,
                                                 _stream()))
            contiguous = own.numel() > 0 and int(own[-1].item()) - int(own[0].item()) + 1 == own.numel()
            first = int(own[0].item()) if contiguous else -1
            self._row_block_cache = cache = (key, indptr, int(nnz.value), own, cols, first)
        _, indptr, nnz, own, cols, first = cache
        if first >= 0:
            # rows of consecutive nodes are consecutive in V: the block is V[start : start + nnz]
            if not hasattr(self, "_nbr_start"):
                self._nbr_start = {}
            if (nvar, first) not in self._nbr_start:
                ind = self.sparsity_pattern(nvar)[1]
                self._nbr_start[(nvar, first)] = int(ind[first * nvar].item())
            s0 = self._nbr_start[(nvar, first)]
            return indptr, cols, V[s0:s0 + nnz]
        vals = torch.empty(nnz, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.fl_row_block_emit(self._h, nvar, _ptr(V), _ptr(own), own.numel(), None, _ptr(indptr), None, _ptr(vals), _stream()))
        return indptr, cols, vals


def sfc_order(points, elements):
    """Permutation that sorts the elements along a Morton curve through their centroids (fl_sfc_order), device tensors in/out."""
    lib = _lib.load()
    dev = elements.device
    pts = to_device(points, torch.float64, dev)
    els = to_device(elements, torch.int64, dev)
    perm = torch.empty(els.shape[0], dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.fl_sfc_order(_ptr(pts), _ptr(els), els.shape[0], els.shape[1], pts.shape[1], pts.shape[0], _ptr(perm), _stream()))
    return perm


def _set_timing(self, enabled=True):
    check(self.lib.fl_set_timing(self._h, 1 if enabled else 0))


def _get_timing(self):
    """(element kernel, stiffness scatter/reduction, nodal reduction) milliseconds of the last assembly call."""
    ms = (C.c_float * 3)()
    check(self.lib.fl_get_timing(self._h, ms))
    return float(ms[0]), float(ms[1]), float(ms[2])


def _set_option(self, option, value):
    check(self.lib.fl_set_option(self._h, int(option), int(value)))


AssemblyHandle.set_option = _set_option
AssemblyHandle.set_timing = _set_timing
AssemblyHandle.get_timing = _get_timing


def measure_fp64_peak(use_dmma=False, iters=20000):
    out = C.c_double(0.0)
    check(_lib.load().fl_measure_fp64_peak(1 if use_dmma else 0, iters, C.byref(out)))
    return out.value
