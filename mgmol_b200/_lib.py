"""ctypes binding of libmgmol_b200.so (the C ABI declared in include/mgmol_b200.h).

The library is the product: if it is missing this module raises at import of
`lib()`, and every compute entry point returns MGB_ENODEVICE without a GPU --
there is no Python/NumPy/torch fallback path anywhere in this package.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgmol_b200.so")

MGB_F32, MGB_F64 = 0, 1
LAP_4M, LAP_2, LAP_4, LAP_6, LAP_8, LAP_4MP = 0, 1, 2, 3, 4, 10
FD_DEL2_4TH_MEHR, FD_DEL2_2ND, FD_DEL2_4TH, FD_DEL2_6TH, FD_DEL2_8TH = 0, 1, 2, 3, 4
FD_RHS_4TH_MEHR1 = 100

c_void_p, c_int, c_size_t, c_double = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_double)


class MgbGrid(ctypes.Structure):
    """struct mgb_grid (pb::Grid + pb::PEenv data)."""
    _fields_ = [
        ("dim", c_int * 3), ("gdim", c_int * 3), ("ghosts", c_int),
        ("h", c_double * 3), ("bc", c_int * 3), ("nproc", c_int * 3),
        ("coord", c_int * 3),
    ]


class MgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mgmol_b200 error %d: %s" % (code, msg))
        self.code = code


_SIGS = {
    "mgb_last_error": (ctypes.c_char_p, []),
    "mgb_version": (c_int, []),
    "mgb_launch_count": (ctypes.c_ulonglong, []),
    "mgb_device_count": (c_int, []),
    "mgb_malloc": (c_int, [ctypes.POINTER(c_void_p), c_size_t]),
    "mgb_free": (c_int, [c_void_p]),
    "mgb_copy_to_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "mgb_copy_to_host": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "mgb_copy_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "mgb_memset": (c_int, [c_void_p, c_int, c_size_t, c_void_p]),
    "mgb_stream_sync": (c_int, [c_void_p]),
    "mgb_fd_apply": (c_int, [c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p,
                             c_void_p, c_int, c_int, c_void_p]),
    "mgb_hpsi": (c_int, [c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p, c_size_t,
                         c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p,
                         c_void_p]),
    "mgb_apply_b": (c_int, [c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p, c_size_t,
                            c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "mgb_residual": (c_int, [c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p, c_size_t,
                             c_void_p, c_size_t, c_void_p, c_int, c_void_p, c_size_t, c_int,
                             c_void_p, c_void_p]),
    "mgb_hpsi_host": (c_int, [c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p, c_size_t,
                              c_void_p, c_void_p, c_size_t, c_int, c_int]),
    "mgb_hpsi_host_peer": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p,
                                   c_size_t, c_void_p, c_void_p, c_size_t, c_int, c_int]),
    "mgb_host_register": (c_int, [c_void_p, c_size_t]),
    "mgb_host_unregister": (c_int, [c_void_p]),
    "mgb_hpsi_last_path": (c_int, []),
    "mgb_hpsi_last_kernel": (ctypes.c_char_p, []),
    "mgb_hpsi_force_path": (c_int, [c_int]),
    "mgb_gfv_set_with_ghosts": (c_int, [c_int, c_int, ctypes.POINTER(MgbGrid),
                                        c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "mgb_gfv_get_values": (c_int, [c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p,
                                   c_void_p, c_size_t, c_int, c_void_p]),
    "mgb_gfv_trade_boundaries": (c_int, [c_int, ctypes.POINTER(MgbGrid), c_void_p,
                                         c_int, c_void_p]),
    "mgb_gfv_pointwise_product": (c_int, [c_int, ctypes.POINTER(MgbGrid), c_void_p,
                                          c_void_p, c_void_p, c_int, c_void_p]),
    "mgb_axpy": (c_int, [c_int, c_size_t, c_double, c_void_p, c_void_p, c_void_p]),
    "mgb_scal": (c_int, [c_int, c_size_t, c_double, c_void_p, c_void_p]),
    "mgb_dot": (c_int, [c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mgb_dot_cols": (c_int, [c_int, c_size_t, c_int, c_double, c_void_p, c_size_t, c_void_p,
                             c_size_t, c_void_p, c_void_p]),
    "mgb_poisson_solve": (c_int, [c_int, c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p, c_void_p,
                                  c_int, c_int, c_int, c_double, c_int, c_void_p]),
    "mgb_rho_blas3": (c_int, [c_int, c_size_t, c_int, c_void_p, c_size_t, c_void_p, c_int,
                              c_void_p, c_size_t, c_void_p, c_void_p]),
    "mgb_gfv_jacobi": (c_int, [c_int, ctypes.POINTER(MgbGrid), c_void_p, c_void_p,
                               c_void_p, c_int, c_double, c_void_p]),
    "mgb_gfv_restrict3D": (c_int, [c_int, ctypes.POINTER(MgbGrid), c_void_p,
                                   c_void_p, c_int, c_void_p]),
    "mgb_gfv_extend3D": (c_int, [c_int, ctypes.POINTER(MgbGrid), c_void_p,
                                 c_void_p, c_int, c_void_p]),
    "mgb_precond_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int,
                                   ctypes.POINTER(MgbGrid), c_int]),
    "mgb_precond_destroy": (c_int, [c_void_p]),
    "mgb_precond_mg": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_int, c_double,
                               c_void_p]),
    "mgb_precond_set_masks": (c_int, [c_void_p, c_void_p]),
    "mgb_precond_set_comm": (c_int, [c_void_p, c_void_p]),
    "mgb_masks_create": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(MgbGrid),
                                 c_int, c_int, c_int, c_int]),
    "mgb_masks_set": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mgb_masks_commit": (c_int, [c_void_p]),
    "mgb_masks_destroy": (c_int, [c_void_p]),
    "mgb_gfv_app_mask": (c_int, [c_int, c_void_p, c_int, c_int, c_void_p, c_int,
                                 c_void_p]),
    "mgb_app_mask": (c_int, [c_int, c_void_p, c_int, c_void_p, c_size_t, c_int,
                             c_void_p]),
    "mgb_precond_set_mode": (c_int, [c_void_p, c_int]),
    "mgb_precond_last_mode": (c_int, [c_void_p]),
    "mgb_precond_vcycle": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mgb_lap_constants": (c_int, [c_int, ctypes.POINTER(c_double),
                                  ctypes.POINTER(c_double)]),
    "mgb_gamma": (c_double, [c_double, c_int, c_double, c_double]),
    "mgb_gemm_tn": (c_int, [c_int, c_int, c_int, c_size_t, c_double, c_void_p,
                            c_size_t, c_void_p, c_size_t, c_double, c_void_p, c_int,
                            c_void_p]),
    "mgb_syrk_t": (c_int, [c_int, c_int, c_size_t, c_double, c_void_p, c_size_t,
                           c_void_p, c_int, c_void_p]),
    "mgb_set_f32_contraction": (c_int, [c_int]),
    "mgb_debug_tn_plan": (c_int, [c_int, c_int, c_int, c_size_t, c_int, c_int, c_int, c_int,
                                  c_void_p, c_int, ctypes.POINTER(c_int),
                                  ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(c_int),
                                  ctypes.POINTER(c_int)]),
    "mgb_gemm_tn_slabs": (c_int, [c_int, c_int, c_int, c_size_t, c_int, c_double, c_void_p,
                                  c_size_t, c_void_p, c_size_t, c_double, c_void_p, c_int,
                                  c_void_p]),
    "mgb_syrk_t_slabs": (c_int, [c_int, c_int, c_size_t, c_int, c_double, c_void_p, c_size_t,
                                 c_void_p, c_int, c_void_p]),
    "mgb_gemm_nn": (c_int, [c_int, c_size_t, c_int, c_int, c_double, c_void_p,
                            c_size_t, c_void_p, c_int, c_double, c_void_p, c_size_t,
                            c_void_p]),
    "mgb_comm_unique_id": (c_int, [c_void_p]),
    "mgb_comm_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int, c_int]),
    "mgb_comm_destroy": (c_int, [c_void_p]),
    "mgb_allreduce_sum_f64": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "mgb_comm_barrier": (c_int, [c_void_p, c_void_p]),
    "mgb_comm_check": (c_int, [c_void_p]),
    "mgb_peer_register": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mgb_peer_unregister": (c_int, [c_void_p, c_void_p]),
    "mgb_peer_set_color_maps": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "mgb_hpsi_peer": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p,
                              c_size_t, c_void_p, c_void_p, c_size_t, c_int, c_void_p,
                              c_void_p]),
    "mgb_hpsi_timing_report": (None, [c_int]),
    "mgb_kb_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_size_t]),
    "mgb_kb_add_ion": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                               ctypes.POINTER(c_int)]),
    "mgb_kb_commit": (c_int, [c_void_p]),
    "mgb_kb_nrows": (c_int, [c_void_p]),
    "mgb_kb_destroy": (c_int, [c_void_p]),
    "mgb_kb_psi": (c_int, [c_void_p, c_int, c_double, c_void_p, c_size_t, c_int, c_void_p,
                           c_void_p]),
    "mgb_kb_vnlpsi": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_int, c_int,
                              c_void_p]),
    "mgb_hpsi_peer3d": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(MgbGrid), c_void_p,
                                c_size_t, c_void_p, c_void_p, c_void_p, c_size_t, c_int,
                                c_void_p]),
    "mgb_halo_exchange_x": (c_int, [c_void_p, c_int, ctypes.POINTER(MgbGrid), c_int,
                                    c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "mgb_halo_set_color_maps": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "mgb_halo_exchange_ghosted": (c_int, [c_void_p, c_int, ctypes.POINTER(MgbGrid),
                                          c_void_p, c_int, c_void_p]),
}

_lib = None


def exported_symbols():
    """Every symbol include/mgmol_b200.h declares."""
    return sorted(_SIGS)


def lib():
    """The loaded library; raises (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                LIB_PATH + " not found: build it with `python -m mgmol_b200.build` "
                "(there is no fallback implementation)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MgbError(rc, lib().mgb_last_error().decode())
