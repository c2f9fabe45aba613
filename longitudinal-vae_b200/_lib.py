"""ctypes binding of liblvae_b200.so (include/lvae_b200.h).  Fails loudly: no library or no CUDA -> RuntimeError."""
import ctypes as C
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblvae_b200.so")

SPEC_STRIDE = 10
MAX_COMPONENTS = 16
MAX_MASKS = 3
MAX_M = 256
MAX_T = 40
CAT, BIN = 0, 1

EXPORTS = ["lvae_kernel_dense_f64", "lvae_kernel_blocks_f64", "lvae_potrf_batched_f64", "lvae_potri_batched_f64",
           "lvae_kld_stats_stride", "lvae_kld_workspace_doubles", "lvae_kld_head_f64", "lvae_kld_subjects_f64",
           "lvae_kld_tail_f64", "lvae_kld_minibatch_f64", "lvae_ng_step_f64", "lvae_launch_count", "lvae_version",
           "lvae_profile_enable", "lvae_profile_last_ms", "lvae_debug_exp_neg_f64", "lvae_kld_hinv_offset", "lvae_gemm_batched_f64", "lvae_ng_workspace_doubles", "lvae_peer_sum_f64",
           "lvae_kernel_dense_bwd_f64", "lvae_kernel_blocks_bwd_f64", "lvae_kld_head_offsets"]

_dp = C.c_void_p


class KernelSpecT(C.Structure):
    _fields_ = [("n_comp0", C.c_int32), ("n_comp1", C.c_int32), ("n_ls", C.c_int32), ("spec", C.POINTER(C.c_int32))]


class KldProblemT(C.Structure):
    _fields_ = [
        ("L", C.c_int32), ("M", C.c_int32), ("Q", C.c_int32), ("P_b", C.c_int32), ("N_b", C.c_int32),
        ("T_max", C.c_int32), ("sum_T2", C.c_int64), ("natural_gradient", C.c_int32), ("path", C.c_int32),
        ("scale", C.c_double), ("const_term", C.c_double), ("eps", C.c_double), ("ks", KernelSpecT),
        ("x", _dp), ("offsets", _dp), ("mu", _dp), ("log_v", _dp), ("z", _dp), ("m", _dp), ("H", _dp),
        ("lengthscale", _dp), ("outputscale", _dp), ("noise", _dp),
        ("kld_per_latent", _dp), ("grad_m", _dp), ("grad_H", _dp), ("d_mu", _dp), ("d_log_v", _dp),
        ("d_lengthscale", _dp), ("d_outputscale", _dp), ("d_noise", _dp),
        ("stats", _dp), ("workspace", _dp), ("info", _dp),
    ]


_lib = None


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"lvae_b200: {LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    i32, i64, dbl, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
    ksp = C.POINTER(KernelSpecT)
    pp = C.POINTER(KldProblemT)
    lib.lvae_kernel_dense_f64.argtypes = [ksp, i32, i32, i32, i32, i32, vp, i64, i32, vp, i64, i32, vp, vp, vp, vp, vp]
    lib.lvae_kernel_blocks_f64.argtypes = [ksp, i32, i32, i32, i32, vp, vp, i32, i64, vp, vp, vp, vp, vp]
    lib.lvae_kernel_dense_bwd_f64.argtypes = [ksp, i32, i32, i32, i32, i32, vp, i64, i32, vp, i64, i32, vp, vp, vp, vp, vp,
                                              vp, vp]
    lib.lvae_kernel_blocks_bwd_f64.argtypes = [ksp, i32, i32, i32, i32, vp, vp, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.lvae_kernel_dense_bwd_f64.restype = lib.lvae_kernel_blocks_bwd_f64.restype = C.c_int
    lib.lvae_potrf_batched_f64.argtypes = [vp, i32, i64, i32, vp, vp]
    lib.lvae_potri_batched_f64.argtypes = [vp, vp, i32, i64, i32, vp]
    lib.lvae_gemm_batched_f64.argtypes = [i32, i32, i32, i32, i32, dbl, vp, i32, i64, vp, i32, i64, dbl, vp, i32, i64, i32,
                                          i32, vp]
    lib.lvae_gemm_batched_f64.restype = C.c_int
    lib.lvae_kld_stats_stride.argtypes = [i32, i32, i32]
    lib.lvae_kld_stats_stride.restype = i64
    lib.lvae_kld_workspace_doubles.argtypes = [pp]
    lib.lvae_kld_workspace_doubles.restype = i64
    for name in ("lvae_kld_head_f64", "lvae_kld_subjects_f64", "lvae_kld_tail_f64", "lvae_kld_minibatch_f64"):
        getattr(lib, name).argtypes = [pp, vp]
        getattr(lib, name).restype = C.c_int
    lib.lvae_ng_step_f64.argtypes = [vp, vp, vp, vp, vp, dbl, i32, i32, vp, vp, vp]
    lib.lvae_peer_sum_f64.argtypes = [C.POINTER(C.c_uint64), i32, i64, vp, vp]
    lib.lvae_peer_sum_f64.restype = C.c_int
    lib.lvae_ng_workspace_doubles.argtypes = [i32, i32]
    lib.lvae_ng_workspace_doubles.restype = i64
    lib.lvae_kld_hinv_offset.argtypes = [pp]
    lib.lvae_kld_hinv_offset.restype = i64
    lib.lvae_kld_head_offsets.argtypes = [pp, C.POINTER(C.c_int64)]
    lib.lvae_kld_head_offsets.restype = C.c_int
    lib.lvae_debug_exp_neg_f64.argtypes = [vp, vp, i32, vp]
    lib.lvae_debug_exp_neg_f64.restype = C.c_int
    lib.lvae_profile_enable.argtypes = [i32]
    lib.lvae_profile_last_ms.argtypes = [i32]
    lib.lvae_profile_last_ms.restype = C.c_float
    lib.lvae_launch_count.restype = i64
    lib.lvae_version.restype = C.c_char_p
    for name in ("lvae_kernel_dense_f64", "lvae_kernel_blocks_f64", "lvae_potrf_batched_f64",
                 "lvae_potri_batched_f64", "lvae_ng_step_f64"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def require_cuda(*tensors):
    lib = load()
    if not torch.cuda.is_available():
        raise RuntimeError("lvae_b200: CUDA device required (the GP-prior ELBO path has no CPU fallback)")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lvae_b200: tensors must live on a CUDA device")
    return lib


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"lvae_b200: {what}: CUDA error {-rc}")
    raise RuntimeError(f"lvae_b200: {what}: invalid argument (code {rc}: 1=bad argument, 2=size over limit "
                       f"[M<={MAX_M}, T<={MAX_T}], 3=bad kernel spec)")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_spec(structure):
    """KernelSpecT + the numpy table that must stay alive while the struct is in use."""
    table = np.ascontiguousarray(structure.table, dtype=np.int32)
    ks = KernelSpecT(structure.n_comp0, structure.n_comp1, structure.n_ls,
                     table.ctypes.data_as(C.POINTER(C.c_int32)))
    return ks, table
