"""ctypes binding of libflorence_b200.so (the C ABI declared in include/florence_b200.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FL_B200_LIB: path of an alternative build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("FL_B200_LIB") or os.path.join(_HERE, "libflorence_b200.so")

FL_OK = 0
FL_ERR_INVALID, FL_ERR_UNSUPPORTED, FL_ERR_CUDA, FL_ERR_STATE = -1, -2, -3, -4
FL_MODE_COO, FL_MODE_CSR = 0, 1

# every symbol include/florence_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = ["fl_last_error", "fl_version", "fl_create", "fl_destroy", "fl_assemble_explicit", "fl_pattern_build", "fl_pattern_export",
           "fl_pattern_export_data_indices", "fl_assemble_implicit", "fl_assemble_laplacian", "fl_assemble_mass", "fl_explicit_steps",
           "fl_pack_nodes", "fl_unpack_add_nodes", "fl_explicit_update", "fl_measure_fp64_peak", "fl_set_timing", "fl_get_timing", "fl_set_option",
           "fl_dirichlet_build", "fl_dirichlet_export", "fl_dirichlet_apply", "fl_set_contact", "fl_assemble_contact",
           "fl_scatter_nodes", "fl_sum_ordered", "fl_explicit_forces", "fl_gather_pack_nodes", "fl_gather_nodes", "fl_explicit_check",
           "fl_row_block_build", "fl_row_block_emit", "fl_sfc_order"]


class MeshDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("nodeperelem", C.c_int32), ("ngauss", C.c_int32), ("reserved", C.c_int32),
                ("nelem", C.c_int64), ("nnode", C.c_int64), ("points", C.c_void_p), ("elements", C.c_void_p),
                ("bases", C.c_void_p), ("Jm", C.c_void_p), ("AllGauss", C.c_void_p)]


class Material(C.Structure):
    _fields_ = [("material_number", C.c_int32), ("reserved", C.c_int32), ("rho", C.c_double), ("mu", C.c_double), ("mu1", C.c_double),
                ("mu2", C.c_double), ("mu3", C.c_double), ("mue", C.c_double), ("lamb", C.c_double), ("eps_1", C.c_double),
                ("eps_2", C.c_double), ("eps_3", C.c_double), ("eps_e", C.c_double)]


class ExplicitCtrl(C.Structure):
    _fields_ = [("dt", C.c_double), ("fext_scale0", C.c_double), ("fext_scale_step", C.c_double), ("incd_scale0", C.c_double),
                ("incd_scale_step", C.c_double), ("increment", C.c_int64), ("nsteps", C.c_int64)]


class UpdateArgs(C.Structure):
    _fields_ = [("dt", C.c_double), ("fext_scale", C.c_double), ("incd_scale", C.c_double), ("M", C.c_void_p), ("fext", C.c_void_p),
                ("fixed_mask", C.c_void_p), ("inc_dirichlet", C.c_void_p), ("T", C.c_void_p), ("iface_slot", C.c_void_p),
                ("T_iface", C.c_void_p), ("U0", C.c_void_p), ("U00", C.c_void_p), ("Eulerx", C.c_void_p), ("status_dev", C.c_void_p),
                ("growth_keys_dev", C.c_void_p), ("use_element_forces", C.c_int32), ("write_T", C.c_int32)]


class FlorenceB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlorenceB200Error("libflorence_b200.so is not built (run `python -m florence_b200.build`); "
                                "there is no CPU fallback for the assembly hot path")
    lib = C.CDLL(LIB_PATH)
    lib.fl_last_error.restype = C.c_char_p
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    lib.fl_create.argtypes = [C.POINTER(MeshDesc), C.POINTER(vp)]
    lib.fl_destroy.argtypes = [vp]
    lib.fl_assemble_explicit.argtypes = [vp, vp, vp, C.POINTER(Material), i32, vp, vp]
    lib.fl_pattern_build.argtypes = [vp, i32, C.POINTER(i64)]
    lib.fl_pattern_export.argtypes = [vp, i32, vp, vp, vp]
    lib.fl_pattern_export_data_indices.argtypes = [vp, i32, vp, vp, vp]
    lib.fl_dirichlet_build.argtypes = [vp, i32, vp, i64, C.POINTER(i64), C.POINTER(i64)]
    lib.fl_dirichlet_export.argtypes = [vp, vp, vp, vp, vp]
    lib.fl_dirichlet_apply.argtypes = [vp, vp, vp, vp, dbl, vp, vp, vp]
    lib.fl_set_contact.argtypes = [vp, vp, i64, vp, dbl, dbl, dbl, vp]
    lib.fl_assemble_contact.argtypes = [vp, vp, vp, i32, vp]
    lib.fl_assemble_implicit.argtypes = [vp, vp, vp, C.POINTER(Material), i32, i32, i32, vp, vp, vp, vp, vp]
    lib.fl_assemble_laplacian.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    lib.fl_assemble_mass.argtypes = [vp, dbl, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.fl_explicit_steps.argtypes = [vp, C.POINTER(Material), C.POINTER(ExplicitCtrl), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fl_pack_nodes.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.fl_unpack_add_nodes.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.fl_scatter_nodes.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.fl_sum_ordered.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    lib.fl_explicit_update.argtypes = [vp, C.POINTER(UpdateArgs), vp]
    lib.fl_explicit_check.argtypes = [vp, vp, i64, vp, vp]
    lib.fl_explicit_forces.argtypes = [vp, vp, C.POINTER(Material), i64, i64, vp]
    lib.fl_gather_pack_nodes.argtypes = [vp, i32, vp, i64, vp, vp]
    lib.fl_gather_nodes.argtypes = [vp, i32, vp, vp]
    lib.fl_row_block_build.argtypes = [vp, i32, vp, i64, vp, C.POINTER(i64), vp]
    lib.fl_row_block_emit.argtypes = [vp, i32, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.fl_sfc_order.argtypes = [vp, vp, i64, i32, i32, i64, vp, vp]
    lib.fl_set_option.argtypes = [vp, i32, i32]
    lib.fl_set_timing.argtypes = [vp, i32]
    lib.fl_get_timing.argtypes = [vp, C.POINTER(C.c_float)]
    lib.fl_measure_fp64_peak.argtypes = [i32, i32, C.POINTER(dbl)]
    for name in EXPORTS:
        if name != "fl_last_error":
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    """Map a status code to the exception the reference raises in the same situation."""
    if rc == FL_OK:
        return
    msg = load().fl_last_error().decode()
    if rc == FL_ERR_UNSUPPORTED:
        # _LowLevelAssembly_.py:58-60 and _LowLevelAssemblyExplicit_DF_DPF_.pyx:107-109 raise NotImplementedError
        raise NotImplementedError(msg)
    if rc == FL_ERR_INVALID:
        raise ValueError(msg)
    raise FlorenceB200Error("libflorence_b200 error %d: %s" % (rc, msg))
