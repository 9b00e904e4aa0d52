"""ctypes binding of ``libdruglamp_sm100.so`` (C ABI: ``include/druglamp_sm100.h``).

There is no CPU or PyTorch fallback: if the library cannot be loaded, or a call fails, a
``RuntimeError`` is raised.  PyTorch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# DRUGLAMP_LIB: an alternative build of the same ABI (A/B measurements on one box)
LIB_PATH = os.environ.get("DRUGLAMP_LIB") or os.path.join(HERE, "libdruglamp_sm100.so")

DL_F32, DL_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
MUL_NONE, MUL_GELU_GRAD, MUL_RELU_MASK, MUL_VALUE = 0, 1, 2, 3

# fp32 operands: True = 3xTF32 split (fp32-grade, the parity mode), False = plain TF32 (the
# precision the reference itself selects on GPUs via set_float32_matmul_precision, main.py:43)
FP32_PRECISE = True

_lib = None
_I64x3 = C.c_int64 * 3


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p),
        ("preact_out", C.c_void_p), ("mul_aux", C.c_void_p), ("residual", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("batch", _I64x3), ("sa", _I64x3), ("sb", _I64x3), ("sc", _I64x3),
        ("ldr", C.c_int64), ("sr", _I64x3),
        ("drop_seed", C.c_uint64), ("drop_p", C.c_float),
        ("alpha", C.c_float),
        ("dtype_ab", C.c_int32), ("dtype_c", C.c_int32),
        ("trans_a", C.c_int32), ("trans_b", C.c_int32),
        ("act", C.c_int32), ("mul_mode", C.c_int32), ("tile_n", C.c_int32), ("precise", C.c_int32),
        ("split_k", C.c_int32), ("conv_taps", C.c_int32), ("conv_left", C.c_int32),
        ("kred", C.c_int32), ("kred_shift", C.c_int32), ("accumulate", C.c_int32),
        ("colsum_a", C.c_void_p), ("drop_seed_step", C.c_void_p), ("pre_mode", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p), ("lse", C.c_void_p),
        ("raw", C.c_void_p), ("d_o", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("dvec", C.c_void_p),
        ("B", C.c_int64), ("H", C.c_int64), ("S2", C.c_int64), ("Lq", C.c_int64), ("Lk", C.c_int64),
        ("d", C.c_int64),
        ("q_ld", C.c_int64), ("q_sb", C.c_int64), ("q_ss", C.c_int64),
        ("k_ld", C.c_int64), ("k_sb", C.c_int64), ("v_ld", C.c_int64), ("v_sb", C.c_int64),
        ("o_ld", C.c_int64), ("o_sb", C.c_int64), ("o_ss", C.c_int64),
        ("dq_ld", C.c_int64), ("dq_sb", C.c_int64), ("dq_ss", C.c_int64),
        ("dk_ld", C.c_int64), ("dk_sb", C.c_int64), ("dv_ld", C.c_int64), ("dv_sb", C.c_int64),
        ("raw_ld", C.c_int64), ("scale", C.c_float), ("dq_accumulate", C.c_int32),
    ]


class FfnArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w1", C.c_void_p), ("w2", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p),
        ("residual", C.c_void_p), ("hidden", C.c_void_p), ("dact", C.c_void_p), ("y", C.c_void_p),
        ("M", C.c_int64), ("D", C.c_int64), ("Dh", C.c_int64),
        ("ldx", C.c_int64), ("ldh", C.c_int64), ("ldy", C.c_int64), ("ldr", C.c_int64),
        ("drop_p", C.c_float), ("seed1", C.c_uint64), ("seed2", C.c_uint64), ("drop_seed_step", C.c_void_p),
    ]


class SmallLinearArgs(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("pre", C.c_void_p), ("Y", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("ldx", C.c_int64), ("ldw", C.c_int64), ("ldy", C.c_int64),
        ("w_kn", C.c_int32), ("act", C.c_int32), ("bn", C.c_int32), ("training", C.c_int32),
        ("eps", C.c_float), ("momentum", C.c_float),
    ]


_P, _I64, _I32, _F, _U64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_uint64
# name -> argtypes; must list every compute entry point declared in include/druglamp_sm100.h
SIGNATURES = {
    "dl_gemm": [C.POINTER(GemmArgs), _P],
    "dl_attn_fwd": [C.POINTER(AttnArgs), _P],
    "dl_attn_bwd": [C.POINTER(AttnArgs), _P],
    "dl_smallk_mul": [_P, _P, _P, _P, _I64, _I32, _I32, _I64, _P],
    "dl_ffn_fwd": [C.POINTER(FfnArgs), _P],
    "dl_ffn_bwd": [C.POINTER(FfnArgs), _P],
    "dl_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _I64, _I32, _F, _I32, _P],
    "dl_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P],
    "dl_softmax_fwd": [_P, _P, _I64, _I32, _I64, _I32, _P],
    "dl_softmax_bwd": [_P, _P, _P, _I64, _I32, _I64, _F, _I32, _P],
    "dl_colsum": [_P, _P, _I64, _I32, _I64, _I32, _I32, _P],
    "dl_dropout": [_P, _P, _I64, _F, _U64, _P, _I32, _P],
    "dl_act_bwd": [_P, _P, _P, _I64, _I32, _F, _U64, _P, _I32, _P],
    "dl_act_fwd": [_P, _P, _I64, _I32, _I32, _P],
    "dl_l2norm_fwd": [_P, _P, _P, _I64, _I32, _F, _I32, _P],
    "dl_l2norm_bwd": [_P, _P, _P, _P, _I64, _I32, _I32, _P],
    "dl_cast": [_P, _I32, _P, _I32, _I64, _P],
    "dl_copy_rows": [_P, _I64, _P, _I64, _I64, _I64, _P],
    "dl_adamw_step": [_P, _P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _F, _F, _P, _P, _I32, _P],
    "dl_add_pe": [_P, _P, _P, _I64, _I64, _F, _U64, _P, _I32, _P],
    "dl_csr_build": [_P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "dl_spmm_norm": [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P],
    "dl_batchnorm_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _F, _F, _I32, _F, _I32, _P],
    "dl_batchnorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _F, _I32, _P],
    "dl_embed_fill_fwd": [_P, _I32, _P, _P, _P, _I64, _I32, _I32, _I32, _P],
    "dl_embed_fill_bwd": [_P, _I32, _P, _P, _I64, _I32, _I32, _I32, _I32, _P],
    "dl_fillbit_pool": [_P, _P, _P, _P, _I32, _I64, _I32, _I32, _I32, _I32, _P],
    "dl_expand_rows": [_P, _P, _P, _I64, _I32, _I32, _I32, _P],
    "dl_transpose": [_P, _P, _I64, _I32, _I32, _I32, _P],
    "dl_bn_transpose": [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P],
    "dl_site_pool_view_bwd": [_P, _P, _I64, _I32, _I32, _I32, _I32, _P],
    "dl_site_pool_fwd": [_P, _P, _I64, _I32, _I32, _I32, _I64, _I32, _P],
    "dl_site_pool_bwd": [_P, _P, _I64, _I32, _I32, _I32, _I64, _I32, _P],
    "dl_mhla_gate_ln_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _F, _I32, _P],
    "dl_mhla_gate_ln_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P],
    "dl_cm_triplet_fwd": [_P, _P, _I64, _I64, _F, _P, _P, _P],
    "dl_cm_triplet_bwd": [_P, _P, _I64, _I64, _F, _P, _P, _P, _P],
    "dl_cross_entropy_fwd": [_P, _P, _P, _P, _I64, _I32, _I64, _I64, _P, _P, _I32, _P],
    "dl_cross_entropy_bwd": [_P, _P, _P, _P, _I64, _I32, _I64, _I64, _P, _P, _P, _P, _I32, _P],
    "dl_small_linear": [C.POINTER(SmallLinearArgs), _P],
    "dl_head_bn_act_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _I32, _I32, _P],
    "dl_bce_fwd": [_P, _P, _P, _P, _I64, _P],
    "dl_bce_bwd": [_P, _P, _P, _P, _I64, _P],
}


def lib():
    """Load (once) and return the CDLL.  Raises if the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m druglamp_b200.build` "
            "(druglamp_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.dl_version.restype = C.c_int
    L.dl_last_error.restype = C.c_char_p
    L.dl_launch_count.restype = C.c_int64
    for name, sig in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = sig
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().dl_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args, stream_ptr()), name)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dt(t) -> int:
    d = t if isinstance(t, torch.dtype) else t.dtype
    if d == torch.float32:
        return DL_F32
    if d == torch.bfloat16:
        return DL_BF16
    raise TypeError(f"unsupported dtype {d}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("druglamp_b200 kernels need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def launch_count() -> int:
    return int(lib().dl_launch_count())


def _3(v: Sequence[int]):
    v = tuple(int(x) for x in v)
    v = v + (0,) * (3 - len(v))
    return _I64x3(*v)


def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         lda: int, ldb: int, ldc: int, trans_a: bool = False, trans_b: bool = False,
         batch=(1, 1, 1), sa=(0, 0, 0), sb=(0, 0, 0), sc=(0, 0, 0),
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         preact_out: Optional[torch.Tensor] = None, mul_aux: Optional[torch.Tensor] = None,
         mul_mode: int = MUL_NONE, residual: Optional[torch.Tensor] = None, ldr: int = 0,
         sr=(0, 0, 0), drop_p: float = 0.0, drop_seed: int = 0, tile_n: int = 0,
         precise: Optional[bool] = None, split_k: int = 0, conv_taps: int = 0, conv_left: int = 0,
         kred: int = 0, kred_shift: int = 0, accumulate: bool = False,
         colsum_a: Optional[torch.Tensor] = None, pre_mode: int = 0) -> None:
    """Raw strided/batched GEMM (see ``dl_gemm`` in the header); all extents in elements."""
    if A.dtype != B.dtype:
        raise TypeError(f"A and B must share a dtype ({A.dtype} vs {B.dtype})")
    for t in (preact_out, mul_aux, residual):
        if t is not None and t.dtype != out.dtype:
            raise TypeError("epilogue tensors must share C's dtype")
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("bias must be fp32")
    b = tuple(int(x) for x in batch) + (1,) * (3 - len(batch))
    a = GemmArgs(ptr(A), ptr(B), ptr(out), ptr(bias), ptr(preact_out), ptr(mul_aux), ptr(residual),
                 M, N, K, lda, ldb, ldc, _I64x3(*b), _3(sa), _3(sb), _3(sc), ldr, _3(sr),
                 drop_seed, drop_p, alpha, dt(A), dt(out), int(trans_a), int(trans_b), act,
                 mul_mode, tile_n, int(FP32_PRECISE if precise is None else precise), split_k,
                 conv_taps, conv_left, int(kred), kred_shift, int(accumulate), ptr(colsum_a),
                 ptr(DROPOUT_STEP) if drop_p > 0 else None, int(pre_mode))
    if PROFILE is None:
        check(lib().dl_gemm(C.byref(a), stream_ptr()), "dl_gemm")
        return
    check(lib().dl_gemm(C.byref(a), stream_ptr()), "dl_gemm")
    # keep the argument block and its buffers alive so the launch can be re-issued for timing
    PROFILE.append({"flops": 2.0 * M * N * K * b[0] * b[1] * b[2], "args": a,
                    "keep": (A, B, out, bias, preact_out, mul_aux, residual, colsum_a),
                    "shape": (M, N, K, b, int(trans_a), int(trans_b))})


def replay_gemm(rec) -> None:
    """Re-issue a recorded tensor-core launch (dl_gemm, or the fused dl_ffn_fwd / dl_ffn_bwd) on the
    current stream (bench.py roofline pass)."""
    fn = rec.get("fn", "dl_gemm")
    check(getattr(lib(), fn)(C.byref(rec["args"]), stream_ptr()), fn)


# device step counter (int64 scalar tensor) that advances every dropout seed per training step; set
# by kernels.set_dropout_step (train.TrainStep points it at the optimiser's step counter)
DROPOUT_STEP = None

# bench.py sets this to a list to record every dl_gemm launch of one step
PROFILE = None
