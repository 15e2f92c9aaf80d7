"""ctypes binding of `libnbm_b200.so` (the C ABI declared in include/nbm_b200.h).

This is the only way the Python host reaches the CUDA kernels.  There is no CPU fallback: if the
library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NBM_B200_LIB: load another build of the same ABI (A/B timing of kernel variants)
LIB_PATH = os.environ.get("NBM_B200_LIB") or os.path.join(_HERE, "libnbm_b200.so")

c_f = C.c_float
c_fp = C.c_void_p  # all device pointers travel as void*


class NbmError(RuntimeError):
    pass


class Lvl(C.Structure):
    _fields_ = [("phi_g", c_fp), ("xg", c_fp), ("yg", c_fp), ("zg", c_fp),
                ("gx", C.c_int), ("gy", C.c_int), ("gz", C.c_int),
                ("interp", C.c_int), ("perturb_eps", c_f),
                ("corner_phi", c_fp), ("cube_phi", c_fp), ("eval_phi", c_fp)]


class Lattice(C.Structure):
    _fields_ = [("xs", c_fp), ("ys", c_fp), ("zs", c_fp),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("lo", C.c_int * 3), ("hi", C.c_int * 3),
                ("n_shift", C.c_int), ("shift", (c_f * 3) * 7)]


class Assemble(C.Structure):
    _fields_ = [("pts", Lattice),
                ("dx", c_f), ("dy", c_f), ("dz", c_f),
                ("bounds", c_f * 6),
                ("shared", C.c_int), ("site_dims", C.c_int * 3), ("pt_off", C.c_int * 3),
                ("flag", c_fp), ("side", c_fp), ("cidx", c_fp), ("frac", c_fp), ("beta_gamma", c_fp),
                ("mu_m_faces", c_fp), ("mu_p_faces", c_fp),
                ("k_m", c_fp), ("k_p", c_fp), ("f_m", c_fp), ("f_p", c_fp), ("g_dir", c_fp),
                ("w", c_fp), ("rhs", c_fp), ("nl", c_fp), ("irr", c_fp),
                ("n_out", C.c_int64), ("out_stride", C.c_int64 * 3), ("out_off", C.c_int64),
                ("irr_capacity", C.c_int64), ("irr_count", c_fp), ("irr_point", c_fp),
                ("irr_wE", c_fp), ("irr_c", c_fp), ("irr_nl", c_fp), ("irr_nlw", c_fp),
                ("faces", C.c_int), ("cface", c_fp), ("dinv", c_fp), ("irr_wU", c_fp), ("irr_rhs", c_fp),
                ("kv", c_fp), ("coef26", c_fp)]


class Net(C.Structure):
    _fields_ = [("layers_p", C.c_int), ("hidden_p", C.c_int), ("layers_m", C.c_int), ("hidden_m", C.c_int)]


class SharedStep(C.Structure):
    _fields_ = [("net", Net),
                ("nonlinear_m", C.c_int), ("nonlinear_p", C.c_int), ("nl_coef_m", c_f), ("nl_coef_p", c_f),
                ("xe", c_fp), ("ye", c_fp), ("ze", c_fp),
                ("ex", C.c_int), ("ey", C.c_int), ("ez", C.c_int),
                ("side", c_fp), ("w", c_fp), ("rhs", c_fp), ("nl", c_fp),
                ("n_crossed", C.c_int64), ("c_node", c_fp), ("B", c_fp),
                ("n_irr", C.c_int64), ("irr_point", c_fp), ("irr_wE", c_fp), ("irr_c", c_fp),
                ("irr_nl", c_fp), ("irr_nlw", c_fp),
                ("inv_n_points", c_f),
                ("U", c_fp), ("R", c_fp), ("G", c_fp), ("E", c_fp), ("gE", c_fp),
                ("partials", c_fp), ("n_partial_rows", C.c_int), ("loss_grad", c_fp), ("stages", C.c_int),
                ("faces", C.c_int), ("cface", c_fp), ("dinv", c_fp), ("irr_wU", c_fp), ("irr_rhs", c_fp),
                ("kv", c_fp), ("S", c_fp), ("coef26", c_fp), ("pc_params", c_fp), ("pc_d1", C.c_int), ("pc_d2", C.c_int),
                ("pc_scale", c_f), ("n_pc_rows", C.c_int), ("pc_nodes_m", c_fp), ("n_pc_m", C.c_int64), ("pc_nodes_p", c_fp), ("n_pc_p", C.c_int64),
                ("pc_d", c_f * 3),
                ("ge_ptr", c_fp), ("ge_ent", c_fp), ("list_nodes", c_fp), ("n_list", C.c_int64), ("g_ptr", c_fp),
                ("g_ent", c_fp),
                ("stencil_tma", C.c_int), ("Hst", c_fp), ("G2", c_fp), ("Rq", c_fp),
                ("B_soa", c_fp), ("irr_c_soa", c_fp), ("irr_wE_soa", c_fp), ("irr_wU_soa", c_fp)]


class PointsStep(C.Structure):
    _fields_ = [("net", Net),
                ("nonlinear_m", C.c_int), ("nonlinear_p", C.c_int), ("nl_coef_m", c_f), ("nl_coef_p", c_f),
                ("xs", c_fp), ("ys", c_fp), ("zs", c_fp),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("p0", C.c_int64), ("p1", C.c_int64),
                ("dx", c_f), ("dy", c_f), ("dz", c_f),
                ("side", c_fp), ("w", c_fp), ("rhs", c_fp), ("nl", c_fp), ("irr", c_fp),
                ("n_crossed", C.c_int64), ("c_site", c_fp), ("c_pos", c_fp), ("c_cube_side", c_fp), ("B", c_fp),
                ("n_irr", C.c_int64), ("irr_wE", c_fp), ("irr_c", c_fp), ("irr_nl", c_fp), ("irr_nlw", c_fp),
                ("inv_n_points", c_f),
                ("E", c_fp), ("gE", c_fp),
                ("partials", c_fp), ("n_partial_rows", C.c_int), ("loss_grad", c_fp), ("rows", c_fp),
                ("xs7", c_fp), ("ys7", c_fp), ("zs7", c_fp), ("U7", c_fp), ("G7", c_fp),
                ("coef26", c_fp), ("pc_params", c_fp), ("Pc", c_fp), ("pc_d1", C.c_int), ("pc_d2", C.c_int),
                ("pc_scale", c_f), ("n_pc_rows", C.c_int),
                ("xs4", c_fp), ("ys4", c_fp), ("zs4", c_fp), ("side4", c_fp), ("U4", c_fp), ("G4", c_fp),
                ("c_live", c_fp), ("n_live", C.c_int64)]


class Optimizer(C.Structure):
    _fields_ = [("n_params", C.c_int),
                ("lr", c_f), ("decay_rate", c_f), ("transition_steps", c_f), ("max_norm", c_f),
                ("b1", c_f), ("b2", c_f), ("eps", c_f), ("optimizer", C.c_int), ("scheduler", C.c_int)]


# every symbol include/nbm_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "nbm_last_error": (C.c_char_p, []),
    "nbm_version": (C.c_int, []),
    "nbm_ghost_layer_f32": (C.c_int, [c_fp] * 4 + [C.c_int] * 3 + [c_fp] * 4 + [c_fp]),
    "nbm_phi_interp_f32": (C.c_int, [_P(Lvl), c_fp, C.c_int64, c_fp, c_fp]),
    "nbm_classify_f32": (C.c_int, [_P(Lvl), _P(Lattice), c_f, c_f, c_f, c_fp, c_fp, c_fp]),
    "nbm_compact_crossed": (C.c_int, [c_fp, C.c_int64, c_fp, C.c_int64, c_fp, c_fp, c_fp, _P(C.c_size_t), c_fp]),
    "nbm_cutcell_f32": (C.c_int, [_P(Lvl), _P(Lattice), c_f, c_f, c_f, c_fp, C.c_int64, c_fp, c_fp, c_fp, c_fp]),
    "nbm_regression_f32": (C.c_int, [_P(Lvl), _P(Lattice), c_f, c_f, c_f, c_fp, C.c_int64] + [c_fp] * 6 + [c_fp]),
    "nbm_site_weights_f32": (C.c_int, [C.c_int64] + [c_fp] * 10 + [c_fp]),
    "nbm_assemble_f32": (C.c_int, [_P(Assemble), c_fp]),
    "nbm_net_num_params": (C.c_int, [_P(Net)]),
    "nbm_upload_params": (C.c_int, [_P(Net), c_fp, c_fp]),
    "nbm_step_partial_rows": (C.c_int, []),
    "nbm_precond_num_params": (C.c_int, [C.c_int, C.c_int]),
    "nbm_ffma_probe_f32": (C.c_int, [C.c_int, c_fp, _P(C.c_double), c_fp]),
    "nbm_loss_grad_shared_f32": (C.c_int, [_P(SharedStep), c_fp]),
    "nbm_loss_grad_points_f32": (C.c_int, [_P(PointsStep), c_fp]),
    "nbm_comm_set_timeout": (C.c_int, [C.c_double]),
    "nbm_comm_alloc_local": (C.c_int, [_P(C.c_void_p)]),
    "nbm_enable_peer_access": (C.c_int, [C.c_int, C.c_int]),
    "nbm_comm_alloc": (C.c_int, [_P(C.c_void_p), C.c_char_p]),
    "nbm_comm_open_peer": (C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    "nbm_comm_close_peer": (C.c_int, [C.c_void_p]),
    "nbm_comm_free": (C.c_int, [C.c_void_p]),
    "nbm_comm_error": (C.c_int, [C.c_void_p]),
    "nbm_reduce_allreduce_f32": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, _P(C.c_void_p), c_fp, c_fp, c_fp]),
    "nbm_reduce_allreduce_finalize_f32": (C.c_int, [_P(Optimizer), _P(Net), c_fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    _P(C.c_void_p), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "nbm_apply_update_f32": (C.c_int, [_P(Optimizer), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "nbm_finalize_step_f32": (C.c_int, [_P(Optimizer), _P(Net), c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "nbm_upload_staged_params": (C.c_int, [c_fp]),
    "nbm_evaluate_f32": (C.c_int, [_P(Net), _P(Lvl), c_fp, C.c_int64, c_f, c_f, c_f, c_fp, c_fp, c_fp, c_fp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NbmError(
                f"{LIB_PATH} is missing: build it with `python -m jax_dips_b200.build` "
                "(the NBM path has no CPU / eager fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().nbm_last_error().decode("utf-8", "replace")
        raise NbmError(f"{what or 'nbm call'} failed (status {rc}): {msg}")


def ptr(t) -> int:
    """device pointer of a tensor (None -> NULL)"""
    if t is None:
        return None
    assert t.is_contiguous(), "C ABI needs contiguous buffers"
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
