"""
ctypes binding of libbq_b200.so (C ABI: include/bq.h).  No logic lives here: tensors in,
raw device pointers + the current CUDA stream out.  There is no CPU fallback — a missing
library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbq_b200.so")

# bq_kind (include/bq.h)
KIND = {
    "block_fp": 0,
    "block_minifloat": 1,
    "block_log": 2,
    "minifloat_denorm": 3,
    "minifloat_ieee": 4,
    "integer": 5,
    "none": 6,
}
BQ_F32, BQ_BF16 = 0, 1


class BqFormat(Structure):
    _fields_ = [
        ("kind", c_int32),
        ("width", c_int32),
        ("exponent_width", c_int32),
        ("exponent_bias", c_int32),
        ("exponent_bias_width", c_int32),
        ("block_rows", c_int32),
        ("block_cols", c_int32),
        ("fold_zero", c_int32),
    ]


class BqTensor3(Structure):
    _fields_ = [("L", c_int64), ("R", c_int64), ("C", c_int64), ("sL", c_int64), ("sR", c_int64), ("sC", c_int64)]


class BqGemmEpilogue(Structure):
    _fields_ = [("bias", c_void_p), ("residual", c_void_p), ("ldr", c_int64), ("scale", ctypes.c_float), ("act", c_int32),
                ("out_dtype", c_int32), ("qfmt", POINTER(BqFormat)), ("qdir", c_int32), ("n_replicas", c_int32),
                ("replicas", c_void_p * 7)]


class BqIpcHandle(Structure):
    _fields_ = [("reserved", ctypes.c_ubyte * 64), ("offset", c_int64), ("size", c_int64)]


class BqError(RuntimeError):
    pass


_STATUS_EXC = {
    1: ValueError,
    2: NotImplementedError,
    3: ValueError,
    4: ValueError,
    5: RuntimeError,
    6: NotImplementedError,
}

_lib = None


def load():
    """Load libbq_b200.so (built by `__graft_entry__.build()` / csrc/build.sh). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BqError(
            f"{LIB_PATH} not found: the sm_100a extension is not built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.bq_strerror.restype = c_char_p
    lib.bq_strerror.argtypes = [ctypes.c_int]
    lib.bq_last_cuda_error.restype = c_char_p
    lib.bq_abi_version.restype = ctypes.c_int
    lib.bq_quantize_workspace_bytes.restype = c_size_t
    lib.bq_quantize_workspace_bytes.argtypes = [POINTER(BqFormat), POINTER(BqTensor3)]
    lib.bq_quantize.restype = ctypes.c_int
    lib.bq_quantize.argtypes = [POINTER(BqFormat), POINTER(BqTensor3), c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                c_size_t, c_void_p]
    lib.bq_silu_mul_quantize.restype = ctypes.c_int
    lib.bq_silu_mul_quantize.argtypes = [POINTER(BqFormat), POINTER(BqTensor3), c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                         c_size_t, c_void_p]
    lib.bq_gemm_bf16_tn.restype = ctypes.c_int
    lib.bq_gemm_bf16_tn.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int64] * 10 + [c_void_p]
    lib.bq_gemm_bf16_tn_ex.restype = ctypes.c_int
    lib.bq_gemm_bf16_tn_ex.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(BqGemmEpilogue)] + [c_int64] * 6 + [c_void_p]
    lib.bq_gemm_bf16_tn_rope.restype = ctypes.c_int
    lib.bq_gemm_bf16_tn_rope.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(BqFormat), c_int32, c_void_p, c_void_p, c_void_p] + \
        [c_int64] * 9 + [c_void_p]
    lib.bq_gemm_bf16_tn_qkv_rope.restype = ctypes.c_int
    lib.bq_gemm_bf16_tn_qkv_rope.argtypes = [c_void_p] * 6 + [POINTER(BqFormat)] * 3 + [c_void_p] * 3 + [c_int64] * 9 + [c_void_p]
    lib.bq_norm_quantize.restype = ctypes.c_int
    lib.bq_norm_quantize.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, ctypes.c_float, c_int32,
                                     POINTER(BqFormat), POINTER(c_void_p), c_void_p]
    lib.bq_linear_workspace_bytes.restype = c_size_t
    lib.bq_linear_workspace_bytes.argtypes = [POINTER(BqFormat), c_int64, c_int64]
    lib.bq_linear.restype = ctypes.c_int
    lib.bq_linear.argtypes = [POINTER(BqFormat), c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                              c_void_p, c_int64, c_void_p, c_size_t, c_void_p]
    lib.bq_linear_fused.restype = ctypes.c_int
    lib.bq_linear_fused.argtypes = [POINTER(BqFormat), c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                    c_int64, c_void_p]
    lib.bq_packed_weight_bytes.restype = c_size_t
    lib.bq_packed_weight_bytes.argtypes = [POINTER(BqFormat), c_int64, c_int64]
    lib.bq_pack_weight.restype = ctypes.c_int
    lib.bq_pack_weight.argtypes = [POINTER(BqFormat), c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]
    lib.bq_gemm_packed_tn.restype = ctypes.c_int
    lib.bq_gemm_packed_tn.argtypes = [c_void_p, c_void_p, POINTER(BqFormat), c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                      c_int64, c_void_p]
    lib.bq_bmm_workspace_bytes.restype = c_size_t
    lib.bq_bmm_workspace_bytes.argtypes = [POINTER(BqFormat), POINTER(BqFormat), c_int64, c_int64, c_int64, c_int64]
    lib.bq_bmm.restype = ctypes.c_int
    lib.bq_bmm.argtypes = [POINTER(BqFormat), POINTER(BqFormat), c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                           c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.bq_attention_causal.restype = ctypes.c_int
    lib.bq_attention_causal.argtypes = [POINTER(BqFormat), c_void_p, c_void_p, c_void_p, c_void_p] + [c_int64] * 8 + [ctypes.c_float, c_void_p]
    lib.bq_attention_causal_q.restype = ctypes.c_int
    lib.bq_attention_causal_q.argtypes = [POINTER(BqFormat), POINTER(BqFormat), c_void_p, c_void_p, c_void_p, c_void_p] + [c_int64] * 8 + [ctypes.c_float, c_void_p]
    lib.bq_attention_masked.restype = ctypes.c_int
    lib.bq_attention_masked.argtypes = [POINTER(BqFormat), POINTER(BqFormat), c_void_p, c_void_p, c_void_p, c_void_p] + [c_int64] * 8 + [
        ctypes.c_float, c_int32, c_void_p, c_int64, c_void_p]
    lib.bq_split3_bf16.restype = ctypes.c_int
    lib.bq_split3_bf16.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.bq_gemm_split_tn.restype = ctypes.c_int
    lib.bq_gemm_split_tn.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32,
                                     POINTER(c_int32), POINTER(c_int32), c_int64, c_void_p]
    lib.bq_set_attention_precise_exp.restype = None
    lib.bq_set_attention_precise_exp.argtypes = [ctypes.c_int]
    lib.bq_get_attention_precise_exp.restype = ctypes.c_int
    lib.bq_get_attention_precise_exp.argtypes = []
    lib.bq_set_attention_dual_pipeline.restype = None
    lib.bq_set_attention_dual_pipeline.argtypes = [ctypes.c_int]
    lib.bq_get_attention_dual_pipeline.restype = ctypes.c_int
    lib.bq_get_attention_dual_pipeline.argtypes = []
    lib.bq_rope_quantize_split.restype = ctypes.c_int
    lib.bq_rope_quantize_split.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32,
                                           c_int64, c_int64, POINTER(BqFormat), c_void_p, c_void_p, c_void_p]
    lib.bq_split3_bf16_transposed.restype = ctypes.c_int
    lib.bq_split3_bf16_transposed.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_int64, c_void_p]
    lib.bq_bmm_split_tn.restype = ctypes.c_int
    lib.bq_bmm_split_tn.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32,
                                    POINTER(c_int32), POINTER(c_int32), c_int64, c_int64, c_int32, c_void_p]
    lib.bq_set_softmax_smem_rows.restype = None
    lib.bq_set_softmax_smem_rows.argtypes = [ctypes.c_int]
    lib.bq_softmax_quantize.restype = ctypes.c_int
    lib.bq_softmax_quantize.argtypes = [POINTER(BqFormat), c_void_p, c_void_p] + [c_int64] * 8 + [ctypes.c_float, c_int32, c_void_p,
                                                                                                 c_int64, c_void_p]
    lib.bq_rope_quantize.restype = ctypes.c_int
    lib.bq_rope_quantize.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32,
                                     c_int64, c_int64, POINTER(BqFormat), POINTER(BqFormat), c_void_p, c_void_p, c_void_p]
    lib.bq_set_norm_warp_rows.restype = None
    lib.bq_set_norm_warp_rows.argtypes = [ctypes.c_int]
    lib.bq_set_stream_quantizer.restype = None
    lib.bq_set_stream_quantizer.argtypes = [ctypes.c_int]
    lib.bq_set_cta_pairs.restype = None
    lib.bq_set_cta_pairs.argtypes = [ctypes.c_int]
    lib.bq_set_pdl.restype = None
    lib.bq_set_pdl.argtypes = [ctypes.c_int]
    lib.bq_get_pdl.restype = ctypes.c_int
    lib.bq_get_pdl.argtypes = []
    lib.bq_set_small_tiles.restype = None
    lib.bq_set_small_tiles.argtypes = [ctypes.c_int]
    lib.bq_split2_f16_rows.restype = ctypes.c_int
    lib.bq_split2_f16_rows.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]
    lib.bq_gemm_split16_tn.restype = ctypes.c_int
    lib.bq_gemm_split16_tn.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32,
                                       POINTER(c_int32), POINTER(c_int32), c_int64, c_void_p]
    lib.bq_bmm_split16_tn.restype = ctypes.c_int
    lib.bq_bmm_split16_tn.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32,
                                      POINTER(c_int32), POINTER(c_int32), c_int64, c_int64, c_void_p]
    lib.bq_token_ce_workspace_bytes.restype = c_size_t
    lib.bq_token_ce_workspace_bytes.argtypes = [c_int64, c_int64]
    lib.bq_token_ce_mean.restype = ctypes.c_int
    lib.bq_token_ce_mean.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int32, c_int64, c_void_p, c_void_p,
                                     c_size_t, c_void_p]
    lib.bq_ipc_export.restype = ctypes.c_int
    lib.bq_ipc_export.argtypes = [c_void_p, POINTER(BqIpcHandle)]
    lib.bq_ipc_import.restype = ctypes.c_int
    lib.bq_ipc_import.argtypes = [POINTER(BqIpcHandle), POINTER(c_void_p), POINTER(c_void_p)]
    lib.bq_ipc_release.restype = ctypes.c_int
    lib.bq_ipc_release.argtypes = [c_void_p]
    lib.bq_peer_barrier.restype = ctypes.c_int
    lib.bq_peer_barrier.argtypes = [POINTER(c_void_p), c_int32, c_int32, ctypes.c_uint32, c_int32, c_void_p]
    lib.bq_peer_barrier_ex.restype = ctypes.c_int
    lib.bq_peer_barrier_ex.argtypes = [POINTER(c_void_p), c_int32, c_int32, ctypes.c_uint32, c_int32, c_void_p, c_void_p]
    lib.bq_peer_push.restype = ctypes.c_int
    lib.bq_peer_push.argtypes = [c_void_p, POINTER(c_void_p), c_int32, c_int64, c_int64, c_int64, c_int64, c_void_p]
    lib.bq_selftest_log2.restype = ctypes.c_int
    lib.bq_selftest_log2.argtypes = [c_void_p, c_void_p]
    lib.bq_kernel_count.restype = ctypes.c_int
    lib.bq_kernel_name.restype = c_char_p
    lib.bq_kernel_name.argtypes = [ctypes.c_int]
    lib.bq_launch_count.restype = c_int64
    lib.bq_launch_count.argtypes = [ctypes.c_int]
    lib.bq_profile_enable.restype = None
    lib.bq_profile_enable.argtypes = [ctypes.c_int]
    lib.bq_profile_read.restype = ctypes.c_int
    lib.bq_profile_read.argtypes = [ctypes.c_int, POINTER(ctypes.c_double), POINTER(c_int64)]
    # A/B switches from the environment (measurement only; the defaults are the shipped configuration)
    if os.environ.get("BQ_ATTENTION_DUAL_PIPELINE") in ("0", "1"):
        lib.bq_set_attention_dual_pipeline(int(os.environ["BQ_ATTENTION_DUAL_PIPELINE"]))
    if os.environ.get("BQ_ATTENTION_PRECISE_EXP") in ("0", "1"):
        lib.bq_set_attention_precise_exp(int(os.environ["BQ_ATTENTION_PRECISE_EXP"]))
    _lib = lib
    return lib


def check(status: int, what: str):
    if status == 0:
        return
    lib = load()
    msg = lib.bq_strerror(status).decode()
    if status == 5:
        msg += ": " + lib.bq_last_cuda_error().decode()
    raise _STATUS_EXC.get(status, BqError)(f"{what}: {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda_f32(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: llm_mixed_q_b200 runs on CUDA (sm_100a) only — there is no CPU fallback"
        )
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (the reference emulates in fp32), got {t.dtype}")


# A small per-device cache of workspaces (owned by torch's caching allocator).
_ws_cache: dict = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """uint8 scratch of at least nbytes, 256-byte aligned, reused across calls on the same stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def launch_counts() -> dict:
    """kernel name -> launches since the library was loaded."""
    lib = load()
    return {lib.bq_kernel_name(i).decode(): int(lib.bq_launch_count(i)) for i in range(lib.bq_kernel_count())}


def profile_enable(on: bool):
    load().bq_profile_enable(1 if on else 0)


def profile_read() -> dict:
    """kernel name -> (total device ms, launches) for launches recorded since profile_enable(True)."""
    lib = load()
    out = {}
    for i in range(lib.bq_kernel_count()):
        ms, n = ctypes.c_double(0), c_int64(0)
        check(lib.bq_profile_read(i, ctypes.byref(ms), ctypes.byref(n)), "bq_profile_read")
        out[lib.bq_kernel_name(i).decode()] = (ms.value, int(n.value))
    return out
