"""ctypes binding of librlcf_b200.so (the C ABI declared in include/rlcf_b200.h).

The product path has no fallback: if the shared library is missing or a call fails, an exception is raised.
Tensors are passed as raw device pointers (`tensor.data_ptr()`), the stream as torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librlcf_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "rlcf_b200.h")

EPI_F16, EPI_GELU_F16, EPI_RESID_F32, EPI_GELU_BWD_F16, EPI_F32, EPI_ADAMW = range(6)

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    "rlcf_abi_version": [],
    "rlcf_last_error": [],
    "rlcf_launch_count": [],
    "rlcf_set_gemm_cta_group": [_i],
    "rlcf_set_attention_impl": [_i],
    "rlcf_set_gemm_multicast": [_i],
    "rlcf_gemm_f16": [_vp, _i, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "rlcf_gemm_f16_grouped": [_vp, _i, _i64, _vp, _i, _i64, _i, _i, _i, _i, _i, _f, _vp, _i64, _vp, _vp, _vp, _vp, _i,
                              _i64, _vp],
    "rlcf_im2col_f16": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "rlcf_embed_lnpre": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _f, _vp, _vp, _vp],
    "rlcf_embed_text": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "rlcf_layernorm_fwd": [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp],
    "rlcf_layernorm_bwd": [_vp, _i, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp, _i64, _i, _vp, _vp, _i, _i64,
                           _i64, _vp],
    "rlcf_attention_fwd": [_vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "rlcf_attention_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "rlcf_attention_row_fwd": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "rlcf_head_fwd": [_vp, _vp, _i64, _vp, _vp, _i64, _i, _vp, _vp, _f, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp],
    "rlcf_entropy_select": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "rlcf_reward_loss": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "rlcf_reward_loss_multi": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _i, _i,
                               _i, _i, _f, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "rlcf_avg_entropy_loss": [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp],
    "rlcf_avg_entropy_reg": [_vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _vp],
    "rlcf_head_bwd": [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _i,
                      _i64, _i64, _vp],
    "rlcf_head_bwd_ex": [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _f, _vp, _vp, _i, _i, _i, _i,
                         _i, _f, _vp, _vp, _i, _i64, _i64, _vp, _vp, _vp, _vp],
    "rlcf_adamw_step_from": [_vp, _vp, _vp, _vp, _i, _i, _i64, _f, _f, _f, _f, _f, _i, _f, _vp, _i64, _i, _vp],
    "rlcf_transpose_blocks_f16": [_vp, _i, _i, _i, _i, _i, _i, _i64, _vp, _i64, _vp],
    "rlcf_colsum_f16": [_vp, _i, _i, _i, _vp, _i64, _vp],
    "rlcf_seq_sum": [_vp, _i, _i, _i, _i, _vp, _i64, _vp],
    "rlcf_outer_sum": [_vp, _vp, _i, _i, _i, _i, _vp, _i64, _vp],
    "rlcf_embed_prompts": [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp],
    "rlcf_pair_logits": [_vp, _vp, _i64, _i, _i, _i, _i, _f, _vp, _vp],
    "rlcf_ctx_grad": [_vp, _i, _i, _i, _i, _i, _vp, _vp],
    "rlcf_embed_prompts_map": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _vp, _vp],
    "rlcf_vec_grad_map": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "rlcf_adamw_step": [_vp, _vp, _vp, _vp, _i, _i, _i64, _f, _f, _f, _f, _f, _i, _f, _vp, _vp],
    "rlcf_reset_params": [_vp, _vp, _vp, _vp, _i, _i64, _vp],
    "rlcf_cast_f16": [_vp, _i64, _i64, _i64, _vp, _i64, _vp],
    "rlcf_gather_seqs": [_vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _vp],
    "rlcf_transpose_cast_f16": [_vp, _i, _i, _vp, _vp],
    "rlcf_embed_lnpre_sets": [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _f, _vp, _vp, _vp],
    "rlcf_head_fwd_sets": [_vp, _vp, _i64, _vp, _vp, _i64, _i, _vp, _i64, _vp, _f, _i, _i, _i, _i, _f, _vp, _vp, _vp,
                           _vp],
    "rlcf_head_bwd_sets": [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _f, _vp, _vp, _i, _i,
                           _i, _i, _i, _f, _vp, _vp, _i, _i64, _i64, _vp, _vp, _vp, _vp],
    "rlcf_transpose_cast_f16_sets": [_vp, _i, _i, _i, _i64, _vp, _i64, _vp],
    "rlcf_retrieval_loss": [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "rlcf_dfeat_partial": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "rlcf_rowdot": [_vp, _vp, _i, _i, _f, _vp, _i64, _vp],
    "rlcf_adamw_full": [_vp, _vp, _vp, _vp, _i, _i64, _f, _f, _f, _f, _f, _i, _f, _vp, _i64, _i, _vp, _i64, _i64, _i64,
                        _vp],
    "rlcf_transpose_f16_sets": [_vp, _i, _i, _vp, _i, _i64, _vp],
    "rlcf_resample_u8": [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp],
    "rlcf_resample_taps": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "rlcf_bicubic_resize": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "rlcf_augmix_views": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _vp, _vp],
    "rlcf_transpose_blocks_colsum": [_vp, _i, _i, _i, _i, _i, _i64, _vp, _i64, _vp, _i64, _vp],
    "rlcf_gemm_wgrad_adamw": [_vp, _i, _i64, _vp, _i, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i64, _vp, _i64, _i, _vp,
                              _i64, _f, _f, _f, _f, _f, _i, _f, _vp],
    "rlcf_gemm_f32": [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _vp],
    "rlcf_attention_f32": [_vp, _i, _i, _i, _i, _vp, _vp],
    "rlcf_accuracy_count": [_vp, _vp, _i, _i, _vp, _vp],
    "rlcf_add_rows": [_vp, _i64, _vp, _i64, _i, _i64, _vp, _vp],
    "rlcf_scale_rows_exp": [_vp, _vp, _i64, _i, _i, _vp, _vp],
    "rlcf_tied_rows_grad": [_vp, _vp, _i, _i, _i, _vp, _vp, _i64, _vp],
}
_RESTYPES = {"rlcf_last_error": C.c_char_p, "rlcf_launch_count": C.c_uint64}

_lib = None


class RlcfError(RuntimeError):
    """A librlcf_b200 call returned a non-zero status."""


def header_symbols() -> list[str]:
    """Function names declared in include/rlcf_b200.h (used by the CPU-side export test)."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rlcf_[a-z0-9_]+)\s*\(", src)))


def load() -> C.CDLL:
    """Loads the shared library (once). Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RlcfError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C rlcf_b200/csrc`). rlcf_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.rlcf_abi_version() != 1:
        raise RlcfError("librlcf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RlcfError(f"{name} failed (code {rc}): {lib.rlcf_last_error().decode()}")


def launch_count() -> int:
    return int(load().rlcf_launch_count())


def set_attention_impl(impl: int) -> int:
    return int(load().rlcf_set_attention_impl(int(impl)))


def set_gemm_multicast(on: int) -> int:
    return int(load().rlcf_set_gemm_multicast(int(on)))


def set_gemm_cta_group(g: int) -> int:
    return int(load().rlcf_set_gemm_cta_group(int(g)))
