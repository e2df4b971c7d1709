"""ctypes binding of the C ABI in ``include/masr_b200.h`` (libmasr_b200.so, built in-tree by
``__graft_entry__.build()`` / ``masr_b200/build.py``).

There is deliberately no fallback: if the shared library is missing or a call fails, an
exception is raised (BASELINE.json north_star: "no CPU fallback")."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmasr_b200.so")

OK = 0
EPI_BIAS, EPI_BIAS_SILU, EPI_BIAS_RELU, EPI_BIAS_GLU, EPI_BIAS_SCALE, EPI_RESIDUAL = range(6)
STATUS_GAIN_EXCEEDED = 1

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes, exactly the declarations of include/masr_b200.h
SIGNATURES = {
    "masr_abi_version": [],
    "masr_check_device": [],
    "masr_stage_waves_f32": [_vp, _vp, _i, _vp, _vp, _i, _vp],
    "masr_fbank_workspace_bytes": [_i, _i64, C.POINTER(_i64)],
    "masr_wave_gain_f32": [_vp, _vp, _i, _i64, _f, _f, _vp, _vp, _vp, _vp],
    "masr_fbank_f32": [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp],
    "masr_conv1_cmvn_relu_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "masr_conv2_s2_relu_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "masr_gemm_f32": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _f, _vp],
    "masr_gemm_tc_f16x2": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _f, _vp],
    "masr_gemm_tc_residual_ln_f16x2": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                       _i, _i, _i, _f, _vp],
    "masr_gemm_tc_residual_postln_f16x2": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                           _i, _i, _i, _f, _vp],
    "masr_gemm_tc_lnpre_f16x2": [_vp, _i64, _vp, _vp, _f, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _f,
                                 _vp],
    "masr_ctc_head_argmax_tc_f16x2": [_vp, _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _vp, _i64, _vp, _vp, _vp],
    "masr_split_f16": [_vp, _vp, _vp, _i64, _vp],
    "masr_conv1_cmvn_relu_planes_f16": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "masr_conv2_tc_f16x2": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "masr_layernorm_f32": [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _f, _vp],
    "masr_layernorm_split_f16": [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _vp],
    "masr_layernorm2_split_f16": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _vp],
    "masr_layernorm_ada_split_f16": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _vp],
    "masr_affine_split_f16": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp],
    "masr_dwconv_bn_silu_f32": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i, _i, _i, _i, _i, _vp],
    "masr_time_reduce_dw_split_f16": [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp],
    "masr_upsample2_add_f32": [_vp, _vp, _vp, _i64, _i64, _i, _i, _i, _vp],
    "masr_relpos_attention_f32": [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64,
                                  _vp, _vp, _i, _i, _i, _i, _vp],
    "masr_relpos_attention_tc": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64,
                                 _i64, _vp, _vp, _i, _i, _i, _i, _vp],
    "masr_relpos_attention_tc5": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp,
                                  _i64, _i64, _vp, _vp, _i, _i, _i, _i, _vp],
    "masr_dwconv_ln_silu_f32": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i, _i, _i, _i,
                                _i, _f, _vp],
    "masr_dwconv_ln_silu_strided_f32": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i, _i, _i,
                                        _i, _i, _i, _f, _vp],
    "masr_grouped_attention_f32": [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "masr_grouped_attention_cache_f32": [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                         _i, _i, _vp],
    "masr_avgpool2_time_f32": [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _vp],
    "masr_lstm_step_f32": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _i, _i, _i, _i, _vp],
    "masr_lstm_seq_workspace_bytes": [_i, _i, C.POINTER(_i64)],
    "masr_lstm_seq_f32": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _i, _i, _i, _i, _vp, _i64, _vp],
    "masr_stream_append_rows": [_vp, _vp, _i64, _i64, _i, _vp, _vp, _i64, _i64, _vp, _vp, _i, _i, _vp],
    "masr_stream_shift_cache": [_vp, _vp, _i64, _i, _i, _vp, _i, _vp],
    "masr_ctc_frame_argmax_f32": [_vp, _i64, _i, _i, _vp, _vp, _vp, _i64, _vp],
    "masr_ctc_topk_f32": [_vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _vp],
    "masr_ctc_prefix_beam_workspace": [_i, _i, C.POINTER(_i64), C.POINTER(_i64)],
    "masr_ctc_prefix_beam": [_vp, _vp, _vp, _i64, _vp, _i, _i, _i, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp],
    "masr_ctc_prefix_beam_state_size": [C.POINTER(_i64), C.POINTER(_i64)],
    "masr_ctc_prefix_beam_stream": [_vp, _vp, _vp, _i64, _vp, _i, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _i64, _vp, _vp, _vp],
    "masr_ctc_greedy_collapse": [_vp, _vp, _i64, _vp, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp],
}


class MasrB200Error(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library and declare every prototype.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MasrB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). masr_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.masr_last_error.restype = C.c_char_p
    lib.masr_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError here == header/library mismatch
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def last_error() -> str:
    return load().masr_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != OK:
        raise MasrB200Error(f"{what} failed (code {rc}): {last_error()}")


def call(name: str, *args):
    """Invoke an ABI function and raise on a non-zero status."""
    check(getattr(load(), name)(*args), name)
