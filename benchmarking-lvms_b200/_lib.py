"""ctypes binding of libblvm_b200.so (include/blvm_b200.h).  No libtorch linkage: tensors cross the boundary as raw
device pointers + sizes + the current CUDA stream handle.  There is NO fallback: if the library is missing the import
fails, and every op raises if its tensors are not CUDA tensors."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLVM_B200_LIB") or os.path.join(_HERE, "lib", "libblvm_b200.so")  # override = A/B builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C benchmarking-lvms_b200/csrc`) first. blvm_b200 has no CPU/PyTorch fallback."
    )

lib = ctypes.CDLL(LIB_PATH)

_p = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int
_f32 = ctypes.c_float
_f64 = ctypes.c_double

class KLLevelStruct(ctypes.Structure):
    """blvm_kl_level_t (include/blvm_b200.h)."""
    _fields_ = [("mu_q", _p), ("sd_q", _p), ("mu_p", _p), ("sd_p", _p), ("kl", _p), ("lens", _p), ("Tz", _i64), ("Z", _i64),
                ("free_nats", _f64), ("g_mu_q", _p), ("g_sd_q", _p), ("g_mu_p", _p), ("g_sd_p", _p), ("g_kl", _p),
                ("part_kl", _p), ("part_klfn", _p), ("z", _p), ("g_z", _p)]


class ElboStepStruct(ctypes.Structure):
    """blvm_elbo_step_t (include/blvm_b200.h)."""
    _fields_ = [("likelihood", _i32), ("raw_dtype", _i32), ("K", _i32), ("D", _i32), ("num_bins", _i32), ("flags", _i32),
                ("n_levels", _i32), ("rank", _i32), ("world", _i32), ("log_epsilon", _f32), ("B", _i64), ("T", _i64),
                ("y", _p), ("raw", _p), ("x_sl", _p), ("lp_twise", _p), ("graw", _p), ("loss_scale", _p),
                ("gmm_softplus_beta", _f64), ("gmm_sd_add", _f64), ("beta", _f64), ("denom", _f64),
                ("levels", KLLevelStruct * 8), ("workspace", _p), ("sync_counter", _p), ("err_flag", _p),
                ("peer_bases_host", ctypes.POINTER(_p)), ("exchange_counters", _p), ("prev_global_sums", _p),
                ("exchange_err", _p)]


SIGNATURES = {
    "blvm_version": (_i32, []),
    "blvm_last_error_string": (ctypes.c_char_p, []),
    "blvm_dmol_chunks": (_i64, [_i64, _i32, _i32]),
    "blvm_dl_chunks": (_i64, [_i64]),
    "blvm_kl_chunks": (_i64, [_i64]),
    "blvm_dmol_has_fast_path": (_i32, [_i32, _i32]),
    "blvm_set_stream_mode": (_i32, [_i32]),
    "blvm_dmol_fwd": (_i32, [_p, _p, _i32, _p, _i64, _i64, _i32, _i32, _i32, _f32, _i32, _p, _p, _p, _p]),
    "blvm_dmol_fwd_grad": (_i32, [_p, _p, _i32, _p, _p, _f32, _p, _i64, _i64, _i32, _i32, _i32, _f32, _i32, _p, _p, _p, _p, _p]),
    "blvm_gmm_chunks": (_i64, [_i64, _i32, _i32]),
    "blvm_gmm_fwd_grad": (_i32, [_p, _p, _p, _p, _f32, _p, _i64, _i64, _i32, _i32, _i32, _f64, _f64, _f64, _i32, _p, _p, _p, _p]),
    "blvm_gaussian_ll": (_i32, [_p, _p, _p, _p, _i64, _f64, _p, _p, _p, _p]),
    "blvm_dl_fwd_grad": (_i32, [_p, _p, _p, _p, _f32, _i64, _i64, _i32, _f32, _i32, _p, _p, _p, _p, _p]),
    "blvm_kl_gaussian_fwd": (_i32, [_p, _p, _p, _p, _i64, _p, _p]),
    "blvm_kl_gaussian_bwd": (_i32, [_p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p]),
    "blvm_kl_elbo_fwd_grad": (_i32, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _f64, _f32, _p, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "blvm_kl_reduce_fwd_grad": (_i32, [_p, _p, _i64, _i64, _i64, _f64, _f32, _p, _p, _p, _i32, _p]),
    "blvm_kl_elbo_levels_fwd_grad": (_i32, [ctypes.POINTER(KLLevelStruct), _i32, _i64, _f32, _i32, _p]),
    "blvm_kl_elbo_levels_fwd_grad_scaled": (_i32, [ctypes.POINTER(KLLevelStruct), _i32, _i64, _f32, _p, _i32, _p]),
    "blvm_elbo_finalize": (_i32, [_p, _i64, ctypes.POINTER(_p), ctypes.POINTER(_p), ctypes.POINTER(_i64), _i32, _p, _i64, _f64, _f64, _p, _p, _p, _p]),
    "blvm_elbo_finalize_publish": (_i32, [_p, _i64, ctypes.POINTER(_p), ctypes.POINTER(_p), ctypes.POINTER(_i64), _i32, _p, _i64,
                                          _f64, _f64, _p, _p, _p, ctypes.POINTER(_p), _i32, _i32, _p, _p, _p, _p]),
    "blvm_exchange_buffer_bytes": (_i64, []),
    "blvm_elbo_step_workspace_doubles": (_i64, [ctypes.POINTER(ElboStepStruct)]),
    "blvm_elbo_step": (_i32, [ctypes.POINTER(ElboStepStruct), _p]),
    "blvm_last_step_launches": (_i32, []),
    "blvm_row_gate_inplace": (_i32, [_p, _i32, _i64, _i64, _p, _p]),
    "blvm_exchange_consume": (_i32, [_p, _i32, _p, _i32, _f64, _p, _p, _p]),
    "blvm_quantize": (_i32, [_p, _i64, _p, _i64, _p, _p]),
    "blvm_linear_dmol_padded_dim": (_i32, [_i32, _i64]),
    "blvm_linear_dmol_max_ctas": (_i64, []),
    "blvm_linear_dmol_fwd_grad": (_i32, [_p, _p, _p, _p, _i32, _p, _f32, _p, _i64, _i64, _i64, _i32, _i32, _f32, _i32, _p, _p, _p, _i64, _p, _p, _p,
                                         ctypes.POINTER(_i64), _p]),
    "blvm_linear_dmol_reduce_dw": (_i32, [_p, _i64, _i64, _i32, _p, _p, _p]),
    "blvm_dmol_sample_mode": (_i32, [_p, _i32, _i64, _i32, _i32, _f32, ctypes.c_uint64, ctypes.c_uint64, _p, _p, _p, _p]),
    "blvm_scale_inplace": (_i32, [_p, _i64, _p, _p]),
    "blvm_scale_inplace_multi": (_i32, [ctypes.POINTER(_p), ctypes.POINTER(_i64), ctypes.POINTER(_i32), _i32, _p, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args

BLVM_DTYPE_F32, BLVM_DTYPE_F16, BLVM_DTYPE_BF16 = 0, 1, 2
BLVM_FLAG_MASK_OUTPUT = 1
BLVM_FLAG_SKIP_PADDED = 2
BLVM_FLAG_OVERLAP_PREV = 4
BLVM_FLAG_NANSUM_LOSS = 8
BLVM_LIK_NONE, BLVM_LIK_DMOL, BLVM_LIK_DL, BLVM_LIK_GMM = 0, 1, 2, 3
BLVM_MAX_KL_LEVELS = 8
BLVM_KL_TILE = 1024


class BlvmError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.blvm_last_error_string().decode("utf-8", "replace")
        raise BlvmError(f"{what} failed with code {rc}: {msg}")
