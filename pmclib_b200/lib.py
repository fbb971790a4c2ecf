"""ctypes loader for libpmcb200.so.  Fails loudly when the CUDA extension is missing: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMCB200_LIB", os.path.join(HERE, "libpmcb200.so"))  # the override is for A/B kernel experiments


class PmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pmcb200 error {code}: {msg}")
        self.code = code


class StepResultC(C.Structure):
    _fields_ = [("perplexity", C.c_double), ("ess", C.c_double), ("logSum", C.c_double), ("ln_evidence", C.c_double),
                ("maxW", C.c_double), ("maxR", C.c_double), ("nok", C.c_long), ("n_reject", C.c_long),
                ("retry", C.c_int), ("n_alive", C.c_int), ("used_precise_path", C.c_int)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_long, C.c_int)

_lib = None

# every symbol include/pmcb200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "pmcgpu_create", "pmcgpu_destroy", "pmcgpu_last_error", "pmcgpu_version", "pmcgpu_device_count", "pmcgpu_stream", "pmcgpu_set_option", "pmcgpu_get_timing", "pmcgpu_get_stats_timing", "pmcgpu_get_draw_timing",
    "pmcgpu_set_proposal", "pmcgpu_set_band_limit", "pmcgpu_set_target_banana", "pmcgpu_set_target_mix",
    "pmcgpu_set_target_defaults", "pmcgpu_set_target_host", "pmcgpu_set_box", "pmcgpu_resize", "pmcgpu_upload",
    "pmcgpu_download", "pmcgpu_set_host_sink", "pmcgpu_pin_host", "pmcgpu_unpin_host", "pmcgpu_device_arrays", "pmcgpu_draw", "pmcgpu_weights", "pmcgpu_draw_weights", "pmcgpu_weights_host_post",
    "pmcgpu_proposal_log_pdf", "pmcgpu_target_log_pdf", "pmcgpu_normalize", "pmcgpu_perplexity_ess",
    "pmcgpu_filter_histogram", "pmcgpu_set_filter_histogram", "pmcgpu_weighted_mean", "pmcgpu_sample_indices", "pmcgpu_resample_residual", "pmcgpu_clip_weights",
    "pmcgpu_count_indices", "pmcgpu_update_rb", "pmcgpu_update_hard", "pmcgpu_ranindx", "pmcgpu_isinbox",
    "pmcgpu_pmc_step", "pmcgpu_step_given", "pmcgpu_get_proposal", "pmcgpu_shard_range", "pmcgpu_nccl_unique_id",
    "pmcgpu_comm_init", "pmcgpu_comm_set_callback", "pmcgpu_stats_len", "pmcgpu_finalize_update", "pmcgpu_cholesky",
    "pmcgpu_set_target_combine", "pmcgpu_step_given_post", "pmcgpu_comm_allreduce", "pmcgpu_comm_bcast",
    "pmcgpu_comm_gather_store", "pmcgpu_comm_rank", "pmcgpu_target_scalar_product", "pmcgpu_get_info", "pmcgpu_comm_allgather_bytes",
]


class CombineTermC(C.Structure):
    """pmcgpu_combine_term (include/pmcb200.h)."""
    _fields_ = [("kind", C.c_int), ("ndim", C.c_int), ("dim_idx", C.c_void_p), ("b", C.c_double), ("sig1", C.c_double),
                ("K", C.c_int), ("wght", C.c_void_p), ("mean", C.c_void_p), ("L", C.c_void_p), ("detL", C.c_void_p),
                ("df", C.c_void_p)]


def load_library(build_if_missing: bool = False):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from .build import build
            build()
        else:
            raise ImportError(f"{LIB_PATH} is missing: run `python -m pmclib_b200.build` (nvcc, sm_100a). "
                              "pmclib_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dbl, lng, i32 = C.c_void_p, C.c_double, C.c_long, C.c_int
    L.pmcgpu_create.argtypes = [i32, C.POINTER(vp)]
    L.pmcgpu_destroy.argtypes = [vp]
    L.pmcgpu_destroy.restype = None
    L.pmcgpu_last_error.argtypes = [vp]
    L.pmcgpu_last_error.restype = C.c_char_p
    L.pmcgpu_stream.argtypes = [vp]
    L.pmcgpu_stream.restype = vp
    L.pmcgpu_set_option.argtypes = [vp, C.c_char_p, dbl]
    L.pmcgpu_get_timing.argtypes = [vp, C.POINTER(dbl), C.POINTER(lng), C.POINTER(lng)]
    L.pmcgpu_get_stats_timing.argtypes = [vp, C.POINTER(dbl)]
    L.pmcgpu_get_draw_timing.argtypes = [vp, C.POINTER(dbl)]
    L.pmcgpu_set_proposal.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.pmcgpu_set_band_limit.argtypes = [vp, vp]
    L.pmcgpu_set_target_banana.argtypes = [vp, i32, dbl, dbl]
    L.pmcgpu_set_target_mix.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.pmcgpu_set_target_defaults.argtypes = [vp, i32, vp, vp]
    L.pmcgpu_set_target_host.argtypes = [vp, i32]
    L.pmcgpu_set_box.argtypes = [vp, i32, i32, vp]
    L.pmcgpu_resize.argtypes = [vp, lng, i32]
    L.pmcgpu_filter_histogram.argtypes = [vp, i32, i32, i32, lng, C.POINTER(lng)]
    L.pmcgpu_set_filter_histogram.argtypes = [vp, i32, i32, i32]
    L.pmcgpu_weighted_mean.argtypes = [vp, i32, C.POINTER(dbl)]
    L.pmcgpu_clip_weights.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(lng), C.POINTER(dbl)]
    L.pmcgpu_sample_indices.argtypes = [vp, C.c_uint64, C.c_uint32, vp]
    L.pmcgpu_resample_residual.argtypes = [vp, C.c_uint64, C.c_uint32, vp]
    L.pmcgpu_upload.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pmcgpu_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pmcgpu_set_host_sink.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pmcgpu_pin_host.argtypes = [vp, C.c_size_t]
    L.pmcgpu_unpin_host.argtypes = [vp]
    L.pmcgpu_device_arrays.argtypes = [vp, C.POINTER(vp)]
    L.pmcgpu_draw.argtypes = [vp, C.c_uint64, C.c_uint32, lng, C.POINTER(lng)]
    L.pmcgpu_weights.argtypes = [vp, dbl, lng, C.POINTER(lng), C.POINTER(dbl), C.POINTER(dbl)]
    L.pmcgpu_draw_weights.argtypes = [vp, C.c_uint64, C.c_uint32, lng, dbl, C.POINTER(lng), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(lng)]
    L.pmcgpu_weights_host_post.argtypes = [vp, vp, vp, dbl, lng, C.POINTER(lng), C.POINTER(dbl), C.POINTER(dbl)]
    L.pmcgpu_proposal_log_pdf.argtypes = [vp, lng, vp, vp]
    L.pmcgpu_target_log_pdf.argtypes = [vp, lng, vp, vp]
    L.pmcgpu_normalize.argtypes = [vp, dbl, C.POINTER(dbl), C.POINTER(lng)]
    L.pmcgpu_perplexity_ess.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(lng)]
    L.pmcgpu_count_indices.argtypes = [vp, vp]
    L.pmcgpu_update_rb.argtypes = [vp, dbl, lng, C.POINTER(i32), vp, vp, vp, vp, vp]
    L.pmcgpu_update_hard.argtypes = [vp, lng, C.POINTER(i32), vp, vp, vp, vp, vp]
    L.pmcgpu_ranindx.argtypes = [vp, lng, vp, vp]
    L.pmcgpu_isinbox.argtypes = [vp, lng, vp, vp]
    L.pmcgpu_pmc_step.argtypes = [vp, C.c_uint64, C.c_uint32, lng, lng, i32, C.POINTER(i32), C.POINTER(StepResultC),
                                  vp, vp, vp, vp, vp]
    L.pmcgpu_step_given.argtypes = [vp, lng, lng, dbl, i32, C.POINTER(i32), C.POINTER(StepResultC), vp, vp, vp, vp, vp]
    L.pmcgpu_get_proposal.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pmcgpu_shard_range.argtypes = [lng, i32, i32, C.POINTER(lng), C.POINTER(lng)]
    L.pmcgpu_shard_range.restype = None
    L.pmcgpu_nccl_unique_id.argtypes = [vp]
    L.pmcgpu_comm_init.argtypes = [vp, vp, i32, i32]
    L.pmcgpu_comm_set_callback.argtypes = [vp, ALLREDUCE_FN, vp, i32, i32]
    L.pmcgpu_stats_len.argtypes = [i32, i32]
    L.pmcgpu_stats_len.restype = lng
    L.pmcgpu_finalize_update.argtypes = [i32, i32, vp, vp, vp, vp, vp, i32, vp, lng, vp, C.POINTER(i32), vp, vp, vp, vp, vp]
    L.pmcgpu_cholesky.argtypes = [i32, vp, C.POINTER(dbl)]
    L.pmcgpu_set_target_combine.argtypes = [vp, i32, i32, vp]
    L.pmcgpu_step_given_post.argtypes = [vp, vp, vp, lng, lng, dbl, i32, C.POINTER(i32), C.POINTER(StepResultC), vp, vp, vp, vp, vp]
    L.pmcgpu_comm_allreduce.argtypes = [vp, vp, lng, i32]
    L.pmcgpu_comm_bcast.argtypes = [vp, vp, C.c_size_t, i32]
    L.pmcgpu_comm_gather_store.argtypes = [vp, i32, lng, lng, vp, vp, vp, vp, vp]
    L.pmcgpu_comm_rank.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.pmcgpu_target_scalar_product.argtypes = [vp, lng, vp, vp]
    L.pmcgpu_get_info.argtypes = [vp, C.c_char_p, C.POINTER(dbl)]
    L.pmcgpu_comm_allgather_bytes.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t]
    _lib = L
    return L
