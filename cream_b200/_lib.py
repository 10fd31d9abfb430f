"""ctypes binding of libcream_b200.so (the C-ABI drop-in boundary, include/cream_b200.h).

The library is the product: if it is missing the import of any compute entry point fails
loudly — there is no CPU or PyTorch fallback on purpose.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("CREAM_B200_LIB", _HERE / "libcream_b200.so"))

c_void_p, c_int, c_i64, c_float, c_double = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
DTYPE_F32, DTYPE_F16, DTYPE_BF16, DTYPE_F64 = 0, 1, 2, 3
EPI_BF16, EPI_BF16_GELU, EPI_F32_RESID, EPI_BF16_DGELU, EPI_F32_ATOMIC, EPI_F32 = range(6)


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int), ("groups", c_int),
        ("a", c_void_p), ("lda", c_i64), ("a_mn", c_int), ("a_group_off", c_int),
        ("b", c_void_p), ("ldb", c_i64), ("b_mn", c_int), ("b_group_rows", c_i64),
        ("k_groups", c_int), ("k_group_len", c_int),
        ("epi", c_int),
        ("out", c_void_p), ("ldo", c_i64),
        ("out_row_mul", c_int), ("out_g_row", c_int), ("out_g_col", c_int),
        ("aux", c_void_p), ("ldaux", c_i64),
        ("bias", c_void_p),
        ("resid", c_void_p), ("ldr", c_i64),
        ("row_scale", c_void_p), ("rows_per_scale", c_int),
        ("alpha", c_float), ("split_k", c_int), ("cta_pair", c_int),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("B", c_int), ("H", c_int), ("N", c_int), ("head_dim", c_int),
        ("scale", c_float),
        ("qkv", c_void_p), ("ld_qkv", c_i64),
        ("out", c_void_p), ("ld_out", c_i64),
        ("lse", c_void_p),
        ("tk_pack", c_void_p), ("tv_pack", c_void_p), ("tables_per_head", c_int),
        ("idx_a", c_void_p), ("idx_b", c_void_p), ("idx_va", c_void_p), ("idx_vb", c_void_p),
        ("ld_idx", c_int),
        ("bias_pack", c_void_p),
        ("af_grid", c_int), ("af_max_rel", c_int),
        ("dout", c_void_p), ("ld_dout", c_i64),
        ("dqkv", c_void_p), ("ld_dqkv", c_i64),
        ("dtk_pack", c_void_p), ("dtv_pack", c_void_p),
        ("dbias_pack", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_i64),
        ("dense_bias", c_void_p), ("dense_stride_b", c_i64), ("dense_stride_h", c_i64), ("dense_stride_i", c_i64),
        ("ddense", c_void_p),
        ("gp_grid", c_int), ("gp_w", c_int), ("gp_skip_id", c_int),
        ("gp_lut_a", C.c_uint8 * 32), ("gp_lut_b", C.c_uint8 * 32),
        ("causal", c_int),
        ("block_len", c_int),
    ]


VIT_MAX_DEPTH = 32


class VitLayer(C.Structure):
    _fields_ = [
        ("heads", c_int), ("ffn", c_int),
        ("ln1_g", c_void_p), ("ln1_b", c_void_p), ("ln2_g", c_void_p), ("ln2_b", c_void_p),
        ("wqkv", c_void_p), ("bqkv", c_void_p), ("wproj", c_void_p), ("bproj", c_void_p),
        ("wfc1", c_void_p), ("bfc1", c_void_p), ("wfc2", c_void_p), ("bfc2", c_void_p),
        ("tab", c_void_p * 4), ("g_tab", c_void_p * 4),
        ("dp_scale", c_void_p),
        ("g_ln1_g", c_void_p), ("g_ln1_b", c_void_p), ("g_ln2_g", c_void_p), ("g_ln2_b", c_void_p),
        ("g_wqkv", c_void_p), ("g_bqkv", c_void_p), ("g_wproj", c_void_p), ("g_bproj", c_void_p),
        ("g_wfc1", c_void_p), ("g_bfc1", c_void_p), ("g_wfc2", c_void_p), ("g_bfc2", c_void_p),
    ]


class VitDesc(C.Structure):
    _fields_ = [
        ("B", c_int), ("N", c_int), ("E", c_int), ("depth", c_int),
        ("num_classes", c_int), ("in_chans", c_int), ("img_size", c_int), ("patch_size", c_int),
        ("eps", c_float),
        ("pool_first", c_int), ("pool_count", c_int),
        ("qkv_interleaved", c_int), ("qkv_group_rows", c_int),
        ("ld_wqkv", c_i64), ("ld_wproj", c_i64), ("ld_wfc1", c_i64), ("ld_wfc2", c_i64), ("ld_wpatch", c_i64), ("ld_whead", c_i64),
        ("ld_gqkv", c_i64), ("ld_gproj", c_i64), ("ld_gfc1", c_i64), ("ld_gfc2", c_i64), ("ld_gpatch", c_i64), ("ld_ghead", c_i64),
        ("scale", c_float),
        ("af_grid", c_int), ("af_max_rel", c_int),
        ("idx_a", c_void_p), ("idx_b", c_void_p), ("idx_va", c_void_p), ("idx_vb", c_void_p), ("ld_idx", c_int),
        ("gp_grid", c_int), ("gp_w", c_int), ("gp_skip_id", c_int),
        ("gp_lut_a", C.c_uint8 * 32), ("gp_lut_b", C.c_uint8 * 32),
        ("tab_nb", c_int), ("tab_row_off1", c_int),
        ("tab_stride_b", c_i64), ("tab_stride_d", c_i64), ("tabv_stride_b", c_i64), ("tabv_stride_d", c_i64),
        ("images", c_void_p),
        ("wpatch", c_void_p), ("bpatch", c_void_p),
        ("cls", c_void_p), ("pos", c_void_p), ("ld_pos", c_i64),
        ("norm_g", c_void_p), ("norm_b", c_void_p),
        ("whead", c_void_p), ("bhead", c_void_p),
        ("logits", c_void_p), ("ld_logits", c_i64),
        ("dlogits", c_void_p), ("ld_dlogits", c_i64),
        ("g_wpatch", c_void_p), ("g_bpatch", c_void_p), ("g_cls", c_void_p), ("g_pos", c_void_p),
        ("g_norm_g", c_void_p), ("g_norm_b", c_void_p), ("g_whead", c_void_p), ("g_bhead", c_void_p),
        ("arena", c_void_p), ("arena_bytes", c_i64),
        ("layers", VitLayer * VIT_MAX_DEPTH),
    ]


class AdamwSeg(C.Structure):
    _fields_ = [
        ("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p),
        ("shadow", c_void_p),
        ("numel", c_i64),
        ("rows", C.c_int32), ("cols", C.c_int32), ("shadow_ld", C.c_int32), ("qkv_group_rows", C.c_int32),
        ("weight_decay", c_float),
        ("step", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol declared in include/cream_b200.h
SIGNATURES = {
    "cream_version": (C.c_char_p, []),
    "cream_rpe_index_version": (C.c_char_p, []),
    "cream_bind_device": (c_int, [c_int]),
    "cream_rpe_index_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_i64, c_i64, c_i64, c_i64, c_int, c_void_p]),
    "cream_rpe_index_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cream_autoformer_rel_index_host": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "cream_irpe_bucket_ids_host": (c_int, [c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_void_p,
                                           C.POINTER(c_int)]),
    "cream_shadow_cast": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "cream_shadow_qkv": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "cream_gemm_bf16": (c_int, [C.POINTER(GemmDesc), c_void_p]),
    "cream_attn_fwd": (c_int, [C.POINTER(AttnDesc), c_void_p]),
    "cream_attn_bwd": (c_int, [C.POINTER(AttnDesc), c_void_p]),
    "cream_attn_bwd_workspace_bytes": (c_i64, [c_int, c_int, c_int]),
    "cream_layernorm_fwd": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_float, c_void_p, c_i64, c_int,
                                    c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "cream_layernorm_bwd": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_i64, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "cream_layernorm_bwd_cast": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_i64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int,
                                         c_void_p, c_i64, c_void_p, c_int, c_void_p, c_void_p]),
    "cream_patch_im2col": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cream_tokens_assemble_fwd": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_int,
                                          c_int, c_int, c_void_p]),
    "cream_tokens_assemble_bwd": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_int,
                                          c_int, c_int, c_void_p]),
    "cream_pool_fwd": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cream_pool_bwd": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cream_cast_scale": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_int, c_void_p, c_i64, c_int, c_void_p]),
    "cream_bias_grad": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_int, c_void_p]),
    "cream_pack_tables": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_i64, c_i64, c_i64,
                                  c_void_p, c_int, c_int, c_i64, c_i64, c_i64, c_void_p]),
    "cream_pack_tables_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_void_p]),
    "cream_unpack_table_grads_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64,
                                               c_void_p]),
    "cream_vit_arena_bytes": (c_i64, [C.POINTER(VitDesc)]),
    "cream_vit_fwd": (c_int, [C.POINTER(VitDesc), c_void_p]),
    "cream_vit_bwd": (c_int, [C.POINTER(VitDesc), c_int, c_int, c_void_p]),
    "cream_vit_last_launches": (c_int, []),
    "cream_xent_fwd_bwd": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    "cream_adamw_step": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_float, c_float, c_void_p]),
    "cream_adamw_chunk": (c_int, []),
    "cream_unpack_table_grads": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_i64, c_i64, c_i64,
                                         c_void_p, c_int, c_int, c_i64, c_i64, c_i64, c_void_p]),
}

_lib = None


class CreamError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CreamError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(cream_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_ERR = {ERR_ARG: "invalid argument", ERR_CUDA: "CUDA error", ERR_UNSUPPORTED: "unsupported configuration"}


LAUNCHES = [0]   # kernels enqueued through the C ABI (bench.py reports it as gpu_launches)


def check(rc: int, what: str, kernels: int = 1) -> None:
    LAUNCHES[0] += kernels
    if rc != OK:
        raise CreamError(f"{what} failed: {_ERR.get(rc, rc)} (see stderr of libcream_b200)")
