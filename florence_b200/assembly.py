"""Reference-facing plug-in layer: the functions Florence's assembly dispatch looks up, backed by the CUDA library.

Mirrors Florence/FiniteElements/Assembly/_LowLevelAssembly_.py:45-141 (same names, arguments, return tuples and error
behaviour) and the Cython wrappers it calls:
    _LowLevelAssemblyDF__<Material>_ / _LowLevelAssemblyDPF__<Material>_   (_LowLevelAssemblyDF_.pyx:53-160, AOT_Assembler.py)
    _LowLevelAssemblyExplicit_DF_DPF_                                       (_LowLevelAssemblyExplicit_DF_DPF_.pyx:44-152)
    _LowLevelAssemblyPerfectLaplacian_                                      (_LowLevelAssemblyPerfectLaplacian_.pyx)
    ComputeSparsityPattern                                                  (ComputeSparsityPattern.pyx:44-112)
    __TotalConstantMassIntegrand__                                          (_MassIntegrand_.pyx:192-349)
Objects are duck-typed exactly as the reference reads them.  Host arrays go to the device once per call
(state) or once per mesh (cached handle); results come back as numpy / scipy objects, or stay on the device as torch
tensors (DLPack-exportable) when `device_out=True`.
"""
import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import backend
from .backend import AssemblyHandle, MATERIAL_NUMBERS, make_material

__all__ = ['_LowLevelAssembly_', '_LowLevelAssemblyExplicit_', '_LowLevelAssemblyLaplacian_']

has_low_level_dispatcher = True

# material -> constants read by the stamped wrappers (AOT_Assembler.py:109-117, :197-217) and by the explicit wrapper
_CONSTANTS = {
    "LinearElastic": ("mu", "lamb"),
    "IncrementalLinearElastic": ("mu", "lamb"),
    "NeoHookean": ("mu", "lamb"),
    "MooneyRivlin": ("mu1", "mu2", "lamb"),
    "ExplicitMooneyRivlin": ("mu1", "mu2", "lamb"),
    "NearlyIncompressibleMooneyRivlin": (),  # (alpha, beta, kappa), see _material_struct
    "IsotropicElectroMechanics_101": ("mu", "lamb", "eps_1"),
    "IsotropicElectroMechanics_105": ("mu1", "mu2", "lamb", "eps_1", "eps_2"),
    "IsotropicElectroMechanics_108": ("mu1", "mu2", "lamb", "eps_2"),
    "ExplicitIsotropicElectroMechanics_108": ("mu1", "mu2", "lamb", "eps_2"),
}
_IMPLICIT_DF = ("LinearElastic", "IncrementalLinearElastic", "NeoHookean", "MooneyRivlin", "NearlyIncompressibleMooneyRivlin")
_IMPLICIT_DPF = ("IsotropicElectroMechanics_101", "IsotropicElectroMechanics_105", "IsotropicElectroMechanics_108")


def _material_struct(material, name=None):
    name = name or getattr(material, "mtype", type(material).__name__)
    if name not in MATERIAL_NUMBERS or name not in _CONSTANTS:
        raise NotImplementedError("Low level assembly for material {} not available. Consider 'optimise=False' for now".format(name))
    if name == "NearlyIncompressibleMooneyRivlin":
        # the explicit wrapper passes (alpha, beta, kappa) in the (mu1, mu2, mu3) slots (.pyx:82-84)
        consts = dict(mu1=material.alpha, mu2=material.beta, mu3=material.kappa)
    else:
        consts = {k: getattr(material, k) for k in _CONSTANTS[name]}
    rho = getattr(material, "rho", 0.0)
    return make_material(MATERIAL_NUMBERS[name], rho, **consts)


# ------------------------------------------------------------------------------------------------ handle cache
_handle_cache = {}


_FP_SAMPLES = 1 << 16


def _array_key(a):
    """Identity + content fingerprint of a mesh/table array: a mesh that is modified in place (same buffer) must not hit a stale
    device copy.  The fingerprint sums 65 536 strided samples (hashing 100 MB of connectivity in full would cost as much as the
    assembly it guards); code that edits a few entries of mesh.points / mesh.elements in place between assemblies calls
    invalidate_handles(mesh) -- the reference has no such cache because it re-reads the host arrays on every call."""
    if isinstance(a, torch.Tensor):
        flat = a.reshape(-1)
        step = max(1, flat.numel() // _FP_SAMPLES)
        fp = float(flat[::step].double().sum().item()) if flat.numel() else 0.0
        return ("t", a.data_ptr(), tuple(a.shape), fp)
    a = np.asarray(a)
    flat = a.reshape(-1)
    step = max(1, flat.shape[0] // _FP_SAMPLES)
    fp = float(flat[::step].astype(np.float64).sum()) if flat.shape[0] else 0.0
    return ("n", a.__array_interface__["data"][0], a.shape, fp)


def get_handle(mesh, function_space):
    """Device handle for (mesh, function_space); cached so repeated Newton / time steps do not re-upload the mesh.
    A handle owns its scratch buffers: it serves one assembly call at a time (the reference's natives are likewise called with
    the GIL held).  Eviction only drops the cache's reference: an integrator or boundary-condition object that still holds the
    handle keeps it alive, and the device memory is released when the last holder lets go (AssemblyHandle.__del__)."""
    key = (_array_key(mesh.points), _array_key(mesh.elements), _array_key(function_space.Jm))
    ent = _handle_cache.get(key)
    if ent is not None:
        return ent
    if len(_handle_cache) >= 4:
        _handle_cache.pop(next(iter(_handle_cache)))
    h = AssemblyHandle(mesh.points, mesh.elements, function_space.Jm, function_space.AllGauss, getattr(function_space, "Bases", None))
    _handle_cache[key] = h
    return h


def invalidate_handles(mesh=None):
    """Forget the cached device copy of `mesh` (all meshes when None) after an in-place edit of its arrays."""
    if mesh is None:
        _handle_cache.clear()
        return
    pk, ek = _array_key(mesh.points)[:3], _array_key(mesh.elements)[:3]
    for key in [k for k in _handle_cache if k[0][:3] == pk or k[1][:3] == ek]:
        _handle_cache.pop(key)


def clear_handles():
    for h in _handle_cache.values():
        h.close()
    _handle_cache.clear()


# Host results.  The reference returns freshly allocated numpy arrays from every call (np.zeros in the Cython wrappers,
# _LowLevelAssemblyDF_.pyx:90-103), so callers may keep K from several assemblies alive at once (modified Newton, the K/M/D
# combinations of the implicit dynamic integrators).  The default here is the same: every call returns arrays the caller owns.
# Large results (>= 64 MiB) are LEASED pinned buffers: the array handed out is backed by page-locked memory the transfer wrote
# directly, and a weakref finaliser returns the buffer to a free list when the caller's last reference (including views, e.g. the
# csr_matrix built on it) is gone.  A buffer is never written while anything refers to it, so the semantics are those of a fresh
# array, but a loop that drops K before re-assembling pays neither a 2.8 GB allocation with its 700 k first-touch page faults nor a
# host-side copy.  At most _LEASE_MAX buffers per (dtype, size) are out at once; a caller that keeps more results alive gets ordinary
# pageable arrays for the rest (device -> pinned staging buffer -> fresh array, pages populated while the device works).
# `reuse_host_buffers(True)` is the older opt-in: views of a 2-deep ring of pinned buffers, valid until the next-but-one call.
_pinned = {}
_PINNED_RING = 2
_reuse_host_buffers = False


def reuse_host_buffers(enabled=True):
    """Opt in to (or out of) zero-copy host results, see above.  Returns the previous setting."""
    global _reuse_host_buffers
    prev, _reuse_host_buffers = _reuse_host_buffers, bool(enabled)
    return prev


_LEASE_MAX = 3
_lease_free = {}         # (dtype, n) -> [pinned tensors not referenced by any result]
_lease_out = {}          # (dtype, n) -> number of buffers currently backing a live result
_lease_lock = None


def _lease_return(key, buf):
    with _lease_lock:
        _lease_out[key] -= 1
        if key not in _lease_free and len(_lease_free) >= 8:      # results of many different sizes: do not hoard pinned memory
            _lease_free.clear()
        free = _lease_free.setdefault(key, [])
        if len(free) < _LEASE_MAX:
            free.append(buf)


def _lease(t):
    """A pinned buffer for a result of t's size that no live array refers to, or None when _LEASE_MAX are already out."""
    global _lease_lock
    import threading
    if _lease_lock is None:
        _lease_lock = threading.Lock()
    key = (t.dtype, t.numel())
    with _lease_lock:
        free = _lease_free.get(key)
        if free:
            buf = free.pop()
        elif _lease_out.get(key, 0) < _LEASE_MAX:
            buf = None
        else:
            return None, key
        _lease_out[key] = _lease_out.get(key, 0) + 1
    if buf is None:
        try:
            buf = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
        except RuntimeError:                       # no page-locked memory left: fall back to pageable results
            with _lease_lock:
                _lease_out[key] -= 1
            return None, key
    return buf, key


def _leased_array(buf, key):
    """numpy array on a leased buffer; the buffer goes back to the free list when the array and all its views are gone."""
    import weakref
    arr = buf.numpy()
    weakref.finalize(arr, _lease_return, key, buf).atexit = False      # nothing to recycle at interpreter exit
    return arr


def release_host_buffers():
    """Drop the free pinned result buffers (they are re-created on demand)."""
    if _lease_lock is not None:
        with _lease_lock:
            _lease_free.clear()


_BIG = 64 << 20          # bytes; below this a result is simply copied out of its staging buffer
_CHUNK = 64 << 20        # bytes per staging chunk of the pipelined fresh-array path
_stage = {}
_pool = None
_libc = None


def _workers():
    global _pool
    if _pool is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _pool = ThreadPoolExecutor(max(2, min(16, len(os.sched_getaffinity(0)))))
    return _pool


def _populate(addr, nbytes, view):
    """Make the pages of [addr, addr + nbytes) resident and writable: madvise(MADV_POPULATE_WRITE) (Linux >= 5.14; ctypes drops the
    GIL), else one store per page."""
    global _libc
    import ctypes
    if _libc is None:
        try:
            _libc = ctypes.CDLL(None, use_errno=True)
        except OSError:
            _libc = False
    if _libc:
        page = 4096
        lo = (addr + page - 1) & ~(page - 1)
        hi = (addr + nbytes) & ~(page - 1)
        if hi > lo and _libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(hi - lo), 23) == 0:
            return
    view[::max(1, 4096 // view.itemsize)] = 0


def _fresh_begin(n, dtype):
    """Allocate the result array of a large transfer BEFORE the device work is queued and let the worker threads populate its pages in
    the background: the first touch of 2.8 GB (the kernel zeroing 700 k pages) otherwise sits behind the transfer and is what bounds
    it (profiles/host_copy_probe.py: 31-40 GB/s first touch against 57 GB/s for PCIe and 58-72 GB/s for the copy into resident pages)."""
    out = np.empty(n, dtype=dtype)
    pool = _workers()
    parts = pool._max_workers
    step = (n + parts - 1) // parts
    base = out.__array_interface__["data"][0]
    futs = [pool.submit(_populate, base + q * step * out.itemsize, (min((q + 1) * step, n) - q * step) * out.itemsize,
                        out[q * step:min((q + 1) * step, n)]) for q in range(parts) if q * step < n]
    return out, futs


def _fresh_from_device(t, prepared=None):
    """Large result -> freshly allocated numpy array the caller owns: the tensor crosses PCIe in chunks through two pinned
    staging buffers while the host threads copy the previous chunk into the (pageable) result.  `prepared` = (array, futures) of
    _fresh_begin when the caller could allocate the result before the device work."""
    flat = t.reshape(-1)
    n, isz = flat.numel(), flat.element_size()
    npdt = torch.empty(0, dtype=t.dtype).numpy().dtype
    if prepared is not None and prepared[0].shape[0] == n and prepared[0].dtype == npdt:
        out, futs = prepared
    else:
        out, futs = _fresh_begin(n, npdt)
    per = max(1, _CHUNK // isz)
    key = (t.dtype, per)
    if key not in _stage:
        _stage[key] = [torch.empty(per, dtype=t.dtype, pin_memory=True) for _ in range(2)]
    bufs = _stage[key]
    pool = _workers()
    stream = torch.cuda.current_stream()
    pending = [None, None]          # per staging buffer: futures of the host copies still reading it
    events = [None, None]

    def drain(k, lo, hi):
        events[k].synchronize()
        src = bufs[k].numpy()[:hi - lo]
        parts = pool._max_workers
        step = (hi - lo + parts - 1) // parts
        pending[k] = [pool.submit(np.copyto, out[lo + q * step:min(lo + (q + 1) * step, hi)], src[q * step:min((q + 1) * step, hi - lo)])
                      for q in range(parts) if q * step < hi - lo]

    prev = None
    for i, lo in enumerate(range(0, n, per)):
        hi, k = min(lo + per, n), i & 1
        if pending[k] is not None:
            for f in pending[k]:
                f.result()
        bufs[k][:hi - lo].copy_(flat[lo:hi], non_blocking=True)
        events[k] = torch.cuda.Event()
        events[k].record(stream)
        if prev is not None:
            drain(*prev)
        prev = (k, lo, hi)
    if prev is not None:
        drain(*prev)
    for k in (0, 1):
        if pending[k] is not None:
            for f in pending[k]:
                f.result()
    for f in futs:
        f.result()
    return out


def _to_host(t, tag, defer=False, prepared=None):
    """D2H through pinned host memory; returns a numpy array (see the ownership note above)."""
    n = t.numel()
    if not _reuse_host_buffers and n * t.element_size() >= _BIG:
        buf, key = _lease(t)
        if buf is not None:
            buf.copy_(t.reshape(-1), non_blocking=True)
            arr = _leased_array(buf, key)
            if defer:
                ev = torch.cuda.Event()
                ev.record()
                return arr, ev
            torch.cuda.current_stream().synchronize()
            return arr
        out = _fresh_from_device(t, prepared)
        return (out, None) if defer else out
    key = (tag, t.dtype, n)
    ent = _pinned.get(key)
    if ent is None:
        if len(_pinned) > 16:
            _pinned.clear()
        ent = {"bufs": [None] * _PINNED_RING, "next": 0}
        _pinned[key] = ent
    k = ent["next"]
    ent["next"] = (k + 1) % _PINNED_RING
    if ent["bufs"][k] is None:
        ent["bufs"][k] = torch.empty(n, dtype=t.dtype, pin_memory=True)
    buf = ent["bufs"][k]
    buf.copy_(t.reshape(-1), non_blocking=True)
    if defer:
        ev = torch.cuda.Event()
        ev.record()
        return buf, ev
    torch.cuda.current_stream().synchronize()
    return _finish_host(buf)


def _finish_host(buf):
    if isinstance(buf, np.ndarray):
        return buf
    out = buf.numpy()
    if _reuse_host_buffers and out.nbytes >= _BIG:
        return out
    return out.copy()


def _to_host_many(items, prepared=None):
    """D2H of several results with the copies queued back to back, small ones first: the host-side copy-out of a small result
    (T) runs while the large one (the K values) is still crossing PCIe.  `prepared` = {tag: _fresh_begin(...)}."""
    order = sorted(range(len(items)), key=lambda i: items[i][0].numel())
    prepared = prepared or {}
    pend = {i: _to_host(items[i][0], items[i][1], defer=True, prepared=prepared.get(items[i][1])) for i in order}
    out = [None] * len(items)
    for i in order:
        buf, ev = pend[i]
        if ev is not None:
            ev.synchronize()
        out[i] = _finish_host(buf)
    return tuple(out)


def _state_to_device(h, Eulerx, Eulerp):
    x = backend.to_device(Eulerx, torch.float64, h.device)
    p = None if Eulerp is None else backend.to_device(Eulerp, torch.float64, h.device)
    return x, p


# ------------------------------------------------------------------------------------------------ stamped implicit assemblers
def _implicit(matname, fields, fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp, device_out=False):
    """Body of _LowLevelAssemblyDF_<Material>_ / _LowLevelAssemblyDPF_<Material>_ (.pyx:53-160)."""
    h = get_handle(mesh, function_space)
    mat = _material_struct(material, matname)
    form = 1 if fields == "electro_mechanics" else 0
    x, p = _state_to_device(h, Eulerx, Eulerp)
    update = bool(fem_solver.requires_geometry_update)
    if fem_solver.recompute_sparsity_pattern:
        I, J, V, T = h.assemble_implicit(x, p, mat, form, update, mode="coo")
        if device_out:
            return I, J, V, T
        return _to_host_many([(I, "I"), (J, "J"), (V, "V"), (T, "T")])
    prepared = None
    if not device_out and not _reuse_host_buffers:
        # when no leased buffer is free the caller gets a pageable K: allocate it now and populate its pages while the device works
        nnz = h.build_pattern(h.ndim + form)
        key = (torch.float64, nnz)
        if nnz * 8 >= _BIG and not _lease_free.get(key) and _lease_out.get(key, 0) >= _LEASE_MAX:
            prepared = {"V": _fresh_begin(nnz, np.float64)}
    V, T = h.assemble_implicit(x, p, mat, form, update, mode="csr")
    if device_out:
        return V, T
    return _to_host_many([(V, "V"), (T, "T")], prepared)


def _stamp(matname, fields):
    def assembler(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp, device_out=False):
        return _implicit(matname, fields, fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp, device_out)
    prefix = "_LowLevelAssemblyDPF__" if fields == "electro_mechanics" else "_LowLevelAssemblyDF__"
    assembler.__name__ = prefix + matname + "_"
    return assembler


# the reference looks these names up in the module's globals() (_LowLevelAssembly_.py:47-60)
for _m in _IMPLICIT_DF:
    globals()["_LowLevelAssemblyDF__" + _m + "_"] = _stamp(_m, "mechanics")
for _m in _IMPLICIT_DPF:
    globals()["_LowLevelAssemblyDPF__" + _m + "_"] = _stamp(_m, "electro_mechanics")


def _lookup(formulation, material):
    prefix = "_LowLevelAssemblyDF__"
    if formulation.fields == "electro_mechanics":
        prefix = "_LowLevelAssemblyDPF__"
    assembly_func = prefix + type(material).__name__ + "_"
    if assembly_func not in globals():
        raise NotImplementedError("Turning optimise option on for {} material is not supported yet. "
                                  "Consider 'optimise=False' for now".format(type(material).__name__))
    return globals()[assembly_func]


def _LowLevelAssembly_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp):
    """_LowLevelAssembly_.py:45-83: returns (stiffness csr_matrix, T, F=[], mass=[])."""
    func = _lookup(formulation, material)
    nvar = formulation.nvar
    mesh.ChangeType()
    n = nvar * mesh.points.shape[0]
    if fem_solver.recompute_sparsity_pattern:
        I, J, V, T = func(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
        stiffness = csr_matrix((V, (I, J)), shape=(n, n), dtype=np.float64)
    else:
        V, T = func(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
        stiffness = csr_matrix((V, fem_solver.indices, fem_solver.indptr), shape=(n, n), dtype=np.float64)
    F, mass = [], []
    return stiffness, T, F, mass


def _LowLevelAssembly_Par_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp):
    """_LowLevelAssembly_.py:87-105: the raw tuple of the stamped assembler."""
    return _lookup(formulation, material)(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)


# ------------------------------------------------------------------------------------------------ explicit
def _LowLevelAssemblyExplicit_DF_DPF_(function_space, formulation, mesh, material, Eulerx, Eulerp, device_out=False):
    """_LowLevelAssemblyExplicit_DF_DPF_.pyx:44-152: T (nnode*nvar)."""
    h = get_handle(mesh, function_space)
    mat = _material_struct(material)
    if formulation.fields == "mechanics":
        form = 0
    elif formulation.fields == "electro_mechanics":
        form = 1
    else:
        raise NotImplementedError("Explicit low level assembly for {} is not available".format(formulation.fields))
    x, p = _state_to_device(h, Eulerx, Eulerp)
    T = h.assemble_explicit(x, p, mat, form)
    return T if device_out else _to_host(T, "Te")


def _LowLevelAssemblyExplicit_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp):
    """_LowLevelAssembly_.py:109-117."""
    mesh.ChangeType()
    T = _LowLevelAssemblyExplicit_DF_DPF_(function_space, formulation, mesh, material, Eulerx, Eulerp)
    F, mass = [], []
    return T, F, mass


def _LowLevelAssemblyExplicit_Par_(function_space, formulation, mesh, material, Eulerx, Eulerp):
    """_LowLevelAssembly_.py:120-122."""
    return _LowLevelAssemblyExplicit_DF_DPF_(function_space, formulation, mesh, material, Eulerx, Eulerp)


# ------------------------------------------------------------------------------------------------ Laplacian
def _LowLevelAssemblyPerfectLaplacian_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp, device_out=False):
    """_LowLevelAssemblyPerfectLaplacian_.pyx: (I, J, V) or V."""
    ndim = formulation.ndim
    if material.e.shape[0] != ndim:
        raise ValueError("Permittivity tensor has to have a size of (ndim x ndim)")
    h = get_handle(mesh, function_space)
    e = -np.asarray(material.e, dtype=np.float64)                       # .pyx:81
    sym = bool(np.allclose(material.e, material.e.T, atol=1e-8))         # .pyx:82
    if fem_solver.recompute_sparsity_pattern:
        I, J, V = h.assemble_laplacian(e, sym, mode="coo")
        return (I, J, V) if device_out else (_to_host(I, "I"), _to_host(J, "J"), _to_host(V, "V"))
    V = h.assemble_laplacian(e, sym, mode="csr")
    return V if device_out else _to_host(V, "V")


def _LowLevelAssemblyLaplacian_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp):
    """_LowLevelAssembly_.py:126-141."""
    mesh.GetNumberOfNodes()
    mesh.ChangeType()
    if fem_solver.recompute_sparsity_pattern:
        I, J, V = _LowLevelAssemblyPerfectLaplacian_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
        stiffness = csr_matrix((V, (I, J)), shape=(mesh.nnode, mesh.nnode), dtype=np.float64)
    else:
        V = _LowLevelAssemblyPerfectLaplacian_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
        stiffness = csr_matrix((V, fem_solver.indices, fem_solver.indptr), shape=(mesh.nnode, mesh.nnode), dtype=np.float64)
    return stiffness, np.zeros(mesh.nnode, np.float64)


# ------------------------------------------------------------------------------------------------ pattern and mass
def ComputeSparsityPattern(mesh, nvar, squeeze_sparsity_pattern=False, function_space=None, device_out=False):
    """ComputeSparsityPattern.pyx:44-112: (indices, indptr[, data_local_indices, data_global_indices]), all int32."""
    if function_space is None:
        # the pattern does not depend on the tables; a 1-point dummy table lets the handle be built from the mesh alone
        class _FS(object):
            pass
        function_space = _FS()
        ndim, npe = mesh.points.shape[1], mesh.elements.shape[1]
        function_space.Jm = np.zeros((ndim, npe, 1))
        function_space.AllGauss = np.ones((1, 1))
        function_space.Bases = np.zeros((npe, 1))
        h = AssemblyHandle(mesh.points, mesh.elements, function_space.Jm, function_space.AllGauss, function_space.Bases)
    else:
        h = get_handle(mesh, function_space)
    out = h.sparsity_pattern(nvar, with_data_indices=not squeeze_sparsity_pattern)
    if device_out:
        return out
    return tuple(t.cpu().numpy() for t in out)


def __TotalConstantMassIntegrand__(mesh, function_space, formulation, mass_type="lumped", recompute_sparsity_pattern=True,
                                   squeeze_sparsity_pattern=False, indices=None, indptr=None, data_global_indices=None,
                                   data_local_indices=None, rho=None, device_out=False):
    """_MassIntegrand_.pyx:192-349.  rho is read from formulation.constant_mass_integrand's material in the reference; here it is
    passed explicitly or taken from formulation.rho / formulation.material.rho."""
    if rho is None:
        rho = getattr(formulation, "rho", None)
        if rho is None:
            rho = formulation.material.rho
    h = get_handle(mesh, function_space)
    nvar = formulation.nvar
    if mass_type == "lumped":
        M = h.assemble_mass(rho, nvar, "lumped")
        M = M if device_out else M.cpu().numpy()[:, None]
        dummy = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float64)
        return (M,) + dummy if recompute_sparsity_pattern else (M, dummy[2])
    if recompute_sparsity_pattern:
        I, J, V = h.assemble_mass(rho, nvar, "consistent", mode="coo")
        if not device_out:
            I, J, V = I.cpu().numpy(), J.cpu().numpy(), V.cpu().numpy()
        return np.zeros((1, 1)), I, J, V
    V = h.assemble_mass(rho, nvar, "consistent", mode="csr")
    return np.zeros((1, 1)), (V if device_out else V.cpu().numpy())


# ------------------------------------------------------------------------------------------------ drop-in installation
def install(florence_module=None):
    """Rebind Florence's low-level assembly entry points to this back end."""
    if florence_module is None:
        import Florence as florence_module
    import sys
    target = sys.modules[florence_module.__name__ + ".FiniteElements.Assembly._LowLevelAssembly_"]
    asm = sys.modules[florence_module.__name__ + ".FiniteElements.Assembly.Assembly"]
    for name in ("_LowLevelAssembly_", "_LowLevelAssembly_Par_", "_LowLevelAssemblyExplicit_", "_LowLevelAssemblyExplicit_Par_",
                 "_LowLevelAssemblyLaplacian_"):
        setattr(target, name, globals()[name])
        if hasattr(asm, name):
            setattr(asm, name, globals()[name])
    # the process-pool launchers and the partitioner they rely on (Assembly.py:79-82, :672-682; FEMSolver.py:1630-1656)
    for name in ("ImplicitParallelLauncher", "ExplicitParallelLauncher"):
        setattr(asm, name, globals()[name])
    fs = sys.modules.get(florence_module.__name__ + ".Solver.FEMSolver")
    if fs is not None and hasattr(fs, "FEMSolver"):
        from . import parallel
        fs.FEMSolver.PartitionMeshForParallelFEM = parallel.PartitionMeshForParallelFEM
    target.has_low_level_dispatcher = True
    return target


# ------------------------------------------------------------------------------------------------ a1 / a2: the callers
def LowLevelAssembly(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp):
    """Assembly.py:37-95: dispatch on formulation.fields, one-off mass for dynamic analyses, returns (K, T[:,None], F, M)."""
    from time import time
    t_assembly = time()
    if not getattr(material, "has_low_level_dispatcher", True):
        raise RuntimeError("Cannot dispatch to low level module since material {} does not support it".format(type(material).__name__))
    if formulation.fields == "electrostatics":
        stiffness, T = _LowLevelAssemblyLaplacian_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
        fem_solver.assembly_time = time() - t_assembly
        return stiffness, T[:, None], None, None
    M = []
    if getattr(fem_solver, "analysis_type", "static") != "static" and getattr(fem_solver, "is_mass_computed", True) is False:
        M = _mass_for(fem_solver, function_space, formulation, mesh, material)
        fem_solver.is_mass_computed = True
    if getattr(fem_solver, "parallel", False):                           # Assembly.py:79-82
        stiffness, T, F, _ = ImplicitParallelLauncher(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
    else:
        stiffness, T, F, _ = _LowLevelAssembly_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
    fem_solver.assembly_time = time() - t_assembly
    return stiffness, T[:, None], F, M


def ImplicitParallelLauncher(*args, **kwargs):
    """Assembly.py:879-1050 (see parallel.py)."""
    from . import parallel
    return parallel.ImplicitParallelLauncher(*args, **kwargs)


def ExplicitParallelLauncher(*args, **kwargs):
    """Assembly.py:1126-1358 (see parallel.py)."""
    from . import parallel
    return parallel.ExplicitParallelLauncher(*args, **kwargs)


def Assemble(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp):
    """Assembly.py:25-34 with has_low_level_dispatcher=True (this back end has no other path)."""
    return LowLevelAssembly(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)


def _mass_for(fem_solver, function_space, formulation, mesh, material):
    mass_type = getattr(fem_solver, "mass_type", "lumped")
    if fem_solver.recompute_sparsity_pattern:
        out = __TotalConstantMassIntegrand__(mesh, function_space, formulation, mass_type, rho=material.rho)
        M = out[0]
        if mass_type == "consistent":
            n = formulation.nvar * mesh.points.shape[0]
            M = csr_matrix((out[3], (out[1], out[2])), shape=(n, n), dtype=np.float64)
        return M
    out = __TotalConstantMassIntegrand__(mesh, function_space, formulation, mass_type, False, fem_solver.squeeze_sparsity_pattern,
                                         fem_solver.indices, fem_solver.indptr, rho=material.rho)
    M = out[0]
    if mass_type == "consistent":
        n = formulation.nvar * mesh.points.shape[0]
        M = csr_matrix((out[1], fem_solver.indices, fem_solver.indptr), shape=(n, n), dtype=np.float64)
    return M


def AssembleExplicit(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp):
    """Assembly.py:664-715: (T[:,None], F, M); the first call also builds the (lumped or consistent) mass."""
    if not getattr(material, "has_low_level_dispatcher", True):
        raise RuntimeError("Cannot dispatch to low level module, since material {} does not support it".format(type(material).__name__))
    if getattr(fem_solver, "parallel", False):                           # Assembly.py:672-674, :681-682
        T, F, M = ExplicitParallelLauncher(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp), [], []
    else:
        T, F, M = _LowLevelAssemblyExplicit_(fem_solver, function_space, formulation, mesh, material, Eulerx, Eulerp)
    if getattr(fem_solver, "is_mass_computed", True) is True:
        return T[:, None], F, M
    M = _mass_for(fem_solver, function_space, formulation, mesh, material)
    fem_solver.is_mass_computed = True
    return T[:, None], [], M
