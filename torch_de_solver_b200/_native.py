"""ctypes binding of libtedeous_b200.so (include/tdb200.h).  Loading fails loudly: there is no fallback."""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .plan import SEGMENT_DTYPE, TERM_DTYPE, FACTOR_DTYPE, MAX_LAYERS

_lib = None


class NetDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('widths', C.c_int32 * (MAX_LAYERS + 1)), ('n_cparams', C.c_int32)]


class MatField(C.Structure):
    _fields_ = [('var', C.c_int32), ('axis', C.c_int32), ('order', C.c_int32), ('half_width', C.c_int32),
                ('n_edge', C.c_int32), ('coef_off', C.c_int32)]


class MatDesc(C.Structure):
    _fields_ = [('n_eq', C.c_int32), ('n_var', C.c_int32), ('n0', C.c_int32), ('n1', C.c_int32),
                ('n_fields', C.c_int32)]


MAT_BC_DTYPE = np.dtype([('n_rows', '<i8'), ('cell_off', '<i8'), ('tgt_off', '<i8'), ('K', '<i4'), ('var', '<i4'),
                         ('slot', '<i4'), ('term_begin', '<i4'), ('term_end', '<i4'), ('sign', '<f4', (4,))],
                        align=True)

EXPORTS = [
    'tdb200_plan_create', 'tdb200_plan_set_points', 'tdb200_plan_set_slots', 'tdb200_plan_set_impl', 'tdb200_plan_set_row_weights', 'tdb200_plan_set_field_seeds',
    'tdb200_plan_out_size', 'tdb200_plan_n_params', 'tdb200_plan_n_fields', 'tdb200_plan_launches_per_call', 'tdb200_plan_kernel_path', 'tdb200_comm_unique_id', 'tdb200_comm_create', 'tdb200_comm_destroy', 'tdb200_plan_set_comm',
    'tdb200_loss_grad', 'tdb200_eval_fields', 'tdb200_plan_n_params_pad', 'tdb200_jacobian_rows', 'tdb200_plan_destroy',
    'tdb200_mat_plan_create', 'tdb200_mat_plan_set_coeffs', 'tdb200_mat_plan_set_bcs', 'tdb200_mat_loss_grad', 'tdb200_mat_eval_fields',
    'tdb200_mat_plan_out_size', 'tdb200_mat_plan_launches_per_call', 'tdb200_mat_plan_kernel_kind', 'tdb200_mat_plan_set_row_window',
    'tdb200_mat_plan_set_timing', 'tdb200_mat_plan_stencil_ms', 'tdb200_mat_time_stencil',
    'tdb200_mat_plan_destroy', 'tdb200_optimizer_step',
    'tdb200_peer_create', 'tdb200_peer_handle', 'tdb200_peer_open', 'tdb200_peer_halo', 'tdb200_peer_allreduce', 'tdb200_peer_allreduce_vec', 'tdb200_plan_set_peer', 'tdb200_mat_plan_set_peer',
    'tdb200_peer_error', 'tdb200_peer_destroy',
    'tdb200_last_error', 'tdb200_version',
]


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources are newer) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.tdb200_last_error.restype = C.c_char_p
    lib.tdb200_plan_create.argtypes = [C.POINTER(NetDesc), i32, vp, i32, vp, i32, vp, i32, vp, i32, i32,
                                       C.POINTER(vp)]
    lib.tdb200_plan_set_points.argtypes = [vp, vp, i64, vp, i64, vp, i64]
    lib.tdb200_plan_set_slots.argtypes = [vp, vp, vp]
    lib.tdb200_plan_set_impl.argtypes = [vp, i32]
    lib.tdb200_plan_set_row_weights.argtypes = [vp, vp]
    lib.tdb200_plan_set_field_seeds.argtypes = [vp, vp]
    for name in ('tdb200_plan_out_size', 'tdb200_plan_n_params', 'tdb200_plan_n_fields', 'tdb200_mat_plan_out_size',
                 'tdb200_plan_n_params_pad'):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = i64
    lib.tdb200_plan_launches_per_call.argtypes = [vp]
    lib.tdb200_plan_kernel_path.argtypes = [vp]
    lib.tdb200_comm_unique_id.argtypes = [vp]
    lib.tdb200_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    lib.tdb200_comm_destroy.argtypes = [vp]
    lib.tdb200_comm_destroy.restype = None
    lib.tdb200_plan_set_comm.argtypes = [vp, vp]
    lib.tdb200_mat_plan_launches_per_call.argtypes = [vp]
    lib.tdb200_mat_plan_kernel_kind.argtypes = [vp]
    lib.tdb200_mat_plan_set_timing.argtypes = [vp, i32]
    lib.tdb200_mat_plan_stencil_ms.argtypes = [vp, vp]
    lib.tdb200_mat_time_stencil.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.tdb200_mat_plan_set_row_window.argtypes = [vp, i32, i32]
    lib.tdb200_loss_grad.argtypes = [vp, vp, vp, vp]
    lib.tdb200_eval_fields.argtypes = [vp, vp, vp, vp, vp]
    lib.tdb200_jacobian_rows.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.tdb200_plan_destroy.argtypes = [vp]
    lib.tdb200_plan_destroy.restype = None
    lib.tdb200_mat_plan_create.argtypes = [C.POINTER(MatDesc), vp, i32, vp, vp, vp, i32, vp, i32, vp, i32,
                                           C.POINTER(vp)]
    lib.tdb200_mat_plan_set_coeffs.argtypes = [vp, vp, i64]
    lib.tdb200_mat_plan_set_bcs.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp]
    lib.tdb200_mat_loss_grad.argtypes = [vp, vp, vp, vp, vp]
    lib.tdb200_mat_eval_fields.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.tdb200_mat_plan_destroy.argtypes = [vp]
    lib.tdb200_mat_plan_destroy.restype = None
    lib.tdb200_peer_create.argtypes = [i32, i32, i64, i64, i32, C.POINTER(vp)]
    lib.tdb200_peer_handle.argtypes = [vp, vp]
    lib.tdb200_peer_open.argtypes = [vp, vp]
    lib.tdb200_peer_halo.argtypes = [vp, vp, i64, i32, i64, i64, i64, i64, i64, vp]
    lib.tdb200_peer_allreduce.argtypes = [vp, vp, i32, vp]
    lib.tdb200_peer_allreduce_vec.argtypes = [vp, vp, i64, vp]
    lib.tdb200_plan_set_peer.argtypes = [vp, vp]
    lib.tdb200_mat_plan_set_peer.argtypes = [vp, vp, vp]
    lib.tdb200_peer_error.argtypes = [vp, vp]
    lib.tdb200_peer_destroy.argtypes = [vp]
    lib.tdb200_peer_destroy.restype = None
    lib.tdb200_optimizer_step.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().tdb200_last_error().decode()
        raise RuntimeError(f'{what} failed ({rc}): {msg}')


def np_ptr(arr: np.ndarray):
    return arr.ctypes.data_as(C.c_void_p)
