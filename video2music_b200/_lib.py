"""ctypes binding of csrc/libv2m_b200.so (the C ABI declared in include/v2m_b200.h).

There is no CPU fallback and no alternative backend: importing the compute ops
without the built extension raises, and every op refuses to run on anything
other than an sm_100 CUDA device.
"""
import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libv2m_b200.so")

F32, BF16 = 0, 1
MAX_DEC_LAYERS = 8

vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64


class Epilogue(C.Structure):
    _fields_ = [("bias", vp), ("residual", vp), ("ldr", i32), ("res_mod", i32), ("row_scale", vp), ("col_vec", vp),
                ("alpha", C.c_float), ("alpha_cols", i32), ("relu", i32), ("residual_bf16", i32),
                ("head_scatter", i32), ("S", i32), ("H", i32), ("dh", i32), ("cap", i32), ("pos0", i32),
                ("part_stride", i64), ("drop_scale", C.c_float), ("drop_thresh", C.c_uint32), ("drop_seed", C.c_uint32),
                ("drop_after_res", i32), ("drop_seed_dev", vp), ("accumulate", i32), ("residual_gate", i32), ("gate_scale", C.c_float),
                ("pad_", i32)]


class Attn(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("o", vp),
                ("q_sb", i64), ("q_sl", i64), ("k_sb", i64), ("k_sl", i64),
                ("v_sb", i64), ("v_sl", i64), ("o_sb", i64), ("o_sl", i64),
                ("B", i32), ("Hq", i32), ("Hkv", i32), ("Lq", i32), ("Lk", i32), ("dh", i32),
                ("causal", i32), ("Er", vp), ("er_len", i32), ("q_scale", C.c_float),
                ("lse", vp), ("p_out", vp), ("drop_scale", C.c_float), ("drop_thresh", C.c_uint32), ("drop_seed", C.c_uint32),
                ("drop_seed_dev", vp), ("lk_dev", vp)]


class AttnBwd(C.Structure):
    _fields_ = ([(n, vp) for n in ("q", "k", "v", "o", "dO", "lse", "Er", "dq", "dk", "dv", "dEr")] +
                [(n, i64) for n in ("q_sb", "q_sl", "k_sb", "k_sl", "v_sb", "v_sl", "o_sb", "o_sl", "do_sb", "do_sl",
                                    "dq_sb", "dq_sl", "dkv_sb", "dkv_sl")] +
                [(n, i32) for n in ("B", "Hq", "Hkv", "Lq", "Lk", "dh", "causal", "er_len", "dtype")] +
                [("q_scale", C.c_float), ("drop_scale", C.c_float), ("drop_thresh", C.c_uint32), ("drop_seed", C.c_uint32),
                ("drop_seed_dev", vp), ("dq_scale", C.c_float)])


class DecLayer(C.Structure):
    _fields_ = [(n, vp) for n in (
        "w_qkv", "b_qkv", "w_so", "b_so", "w_cq", "b_cq", "w_co", "b_co", "w_f1", "b_f1", "w_f2", "b_f2",
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "ln3_g", "ln3_b", "er", "er_sw", "self_k", "self_v", "cross_k", "cross_v")]


class Decode(C.Structure):
    _fields_ = ([("dtype", i32)] +
                [(n, i32) for n in ("B", "H", "E", "FF", "S", "cap", "n_layers", "er_len", "vocab", "vocab_limit",
                                    "primer_len", "chord_embed")] +
                [("layer", DecLayer * MAX_DEC_LAYERS)] +
                [(n, vp) for n in ("lnf_g", "lnf_b", "w_out", "b_out", "emb_root", "emb_attr", "emb_chord",
                                   "w_chord", "wc_key", "b_chord", "pe", "key", "gen", "gen_root", "gen_attr",
                                   "step", "h", "r", "qbuf", "ctx", "ff", "logits", "logits_all", "xn")] +
                [(n, i32) for n in ("sample", "max_conseq_N", "max_conseq_chord", "pad_")] + [("uniforms", vp)])


# every symbol include/v2m_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "v2m_abi_version", "v2m_last_error", "v2m_struct_size", "v2m_device_ok", "v2m_gemm_f32", "v2m_gemm_f32_strided", "v2m_gemm_bf16", "v2m_gemm_bf16_general", "v2m_attn_fwd", "v2m_attn_bwd", "v2m_attn_bwd_tc", "v2m_attn_bwd_tc_workspace", "v2m_dy_prep",
    "v2m_layernorm_bwd", "v2m_embed_bwd", "v2m_amt_loss", "v2m_count_valid", "v2m_amt_metrics", "v2m_amt_correspondence", "v2m_adam_step",
    "v2m_layernorm", "v2m_embed_sum", "v2m_concat_features", "v2m_cast_2d", "v2m_decode_run",
    "v2m_decode_run_stream", "v2m_kv_interleave", "v2m_decode_launches_per_step", "v2m_decode_probe", "v2m_binary_f32", "v2m_rope_quirk", "v2m_mamba_conv_silu", "v2m_mamba_step_conv", "v2m_mamba_step_ssm", "v2m_selective_scan_fwd", "v2m_selective_scan_workspace", "v2m_selective_scan_bwd_workspace", "v2m_selective_scan_bwd", "v2m_mamba_conv_silu_bwd", "v2m_rmsnorm", "v2m_rmsnorm_bwd", "v2m_pscan_fwd", "v2m_pscan_bwd", "v2m_moe_route", "v2m_moe_permute", "v2m_moe_grouped_gemm", "v2m_gemm_bf16_grouped", "v2m_gemm_bf16_kgrouped", "v2m_swiglu_pair_bwd_bf16", "v2m_moe_group_colsum_bf16", "v2m_swiglu_pair_bf16", "v2m_moe_combine", "v2m_moe_combine_bwd", "v2m_swiglu_bwd", "v2m_moe_grouped_dw", "v2m_dw_f32", "v2m_step_linear_f32", "v2m_step_attn_f32",
]

_lib: Optional[C.CDLL] = None
_launches = 0          # kernels launched through this binding (bench.py reports it as gpu_launches)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "video2music_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU / PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.v2m_last_error.restype = C.c_char_p
    lib.v2m_decode_launches_per_step.restype = i64
    lib.v2m_struct_size.restype = i64
    lib.v2m_struct_size.argtypes = [i32]
    for name in EXPORTS:
        getattr(lib, name)          # raises AttributeError if the library is stale
    lib.v2m_gemm_f32.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, C.POINTER(Epilogue), vp]
    lib.v2m_gemm_bf16.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, C.POINTER(Epilogue), vp]
    lib.v2m_attn_fwd.argtypes = [C.POINTER(Attn), i32, vp]
    lib.v2m_gemm_bf16_general.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, i32, C.POINTER(Epilogue), vp]
    lib.v2m_gemm_f32_strided.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, C.POINTER(Epilogue), vp]
    lib.v2m_attn_bwd.argtypes = [C.POINTER(AttnBwd), vp]
    lib.v2m_attn_bwd_tc.argtypes = [C.POINTER(AttnBwd), vp, i64, vp]
    lib.v2m_attn_bwd_tc_workspace.argtypes = [i32, i32, i32, i32, i32]
    lib.v2m_attn_bwd_tc_workspace.restype = i64
    lib.v2m_dy_prep.argtypes = [vp, i32, i64, vp, i32, i64, i32, C.c_float, i32, vp, i32, i64, vp, i32, i32, C.c_float, C.c_uint32, C.c_uint32, vp, vp]
    lib.v2m_layernorm_bwd.argtypes = [vp, i32, vp, vp, i32, vp, i32, vp, vp, i32, i32, C.c_float, vp]
    lib.v2m_embed_bwd.argtypes = [vp, vp, i32, i64, vp, i32, i32, vp]
    lib.v2m_rope_quirk.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.v2m_amt_metrics.argtypes = [vp, vp, i32, i32, i64, i32, i32, i32, vp, vp]
    lib.v2m_amt_correspondence.argtypes = [vp, vp, vp, i32, i32, i32, C.c_float, i32, vp, vp]
    lib.v2m_amt_loss.argtypes = [vp, vp, vp, i32, i32, i64, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]
    lib.v2m_count_valid.argtypes = [vp, i32, i64, vp, vp]
    lib.v2m_adam_step.argtypes = [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i32, C.c_float, vp, vp, i32, vp, vp]
    lib.v2m_layernorm.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, vp, i32, i32, i32, C.c_float, vp]
    lib.v2m_step_linear_f32.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]
    lib.v2m_step_attn_f32.argtypes = [vp, i64, vp, vp, i64, i64, vp, i64, i32, i32, i32, i32, i32, vp, C.c_float, vp]
    lib.v2m_embed_sum.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.v2m_concat_features.argtypes = [vp, i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, vp]
    lib.v2m_cast_2d.argtypes = [vp, i32, i64, vp, i32, i64, i32, i32, i32, vp]
    lib.v2m_binary_f32.argtypes = [vp, vp, vp, i64, i32, C.c_float, vp]
    lib.v2m_mamba_step_conv.argtypes = [vp, i64, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.v2m_mamba_step_ssm.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.v2m_gemm_bf16_kgrouped.argtypes = [vp, i32, vp, i32, vp, i32, i64, i32, i32, i32, i32, vp, vp]
    lib.v2m_swiglu_pair_bwd_bf16.argtypes = [vp, vp, vp, i64, i32, vp]
    lib.v2m_moe_group_colsum_bf16.argtypes = [vp, i64, vp, i32, vp, i32, i32, vp]
    lib.v2m_decode_run.argtypes = [C.POINTER(Decode), i32, i32, vp]
    lib.v2m_decode_run_stream.argtypes = [C.POINTER(Decode), i32, i32, vp, i32, vp]
    lib.v2m_kv_interleave.argtypes = [vp, vp, vp, i64, i32, vp]
    lib.v2m_decode_launches_per_step.argtypes = [C.POINTER(Decode)]
    lib.v2m_decode_probe.argtypes = [C.POINTER(Decode), i32, i32, vp]
    lib.v2m_mamba_conv_silu.argtypes = [vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, vp]
    lib.v2m_selective_scan_fwd.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, vp, i64, vp]
    lib.v2m_selective_scan_workspace.argtypes = [i32, i32, i32, i32]
    lib.v2m_selective_scan_workspace.restype = i64
    lib.v2m_selective_scan_bwd_workspace.argtypes = [i32, i32, i32, i32]
    lib.v2m_selective_scan_bwd_workspace.restype = i64
    lib.v2m_selective_scan_bwd.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, i64,
                                           vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.v2m_mamba_conv_silu_bwd.argtypes = [vp, i64, vp, vp, vp, i64, vp, i64, vp, vp, i32, i32, i32, i32, vp]
    lib.v2m_rmsnorm.argtypes = [vp, vp, vp, i32, i32, C.c_float, vp]
    lib.v2m_rmsnorm_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, vp]
    lib.v2m_pscan_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.v2m_pscan_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.v2m_moe_permute.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, i32, vp]
    lib.v2m_gemm_bf16_grouped.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp]
    lib.v2m_swiglu_pair_bf16.argtypes = [vp, vp, i64, i32, vp]
    lib.v2m_moe_grouped_gemm.argtypes = [vp, i32, vp, vp, vp, vp, i64, i64, vp, i32, i32, vp, i32, i32, i32, vp]
    lib.v2m_moe_combine.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.v2m_moe_combine_bwd.argtypes = [vp, vp, vp, vp, vp, C.c_float, i32, i32, i32, i32, vp, vp, vp]
    lib.v2m_swiglu_bwd.argtypes = [vp, vp, vp, vp, i64, i32, vp]
    lib.v2m_moe_grouped_dw.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp, i32, i32, vp]
    lib.v2m_dw_f32.argtypes = [vp, i32, vp, i32, i32, vp, vp, i32, i32, vp]
    lib.v2m_moe_route.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    for which, cls in enumerate((Epilogue, Attn, DecLayer, Decode, AttnBwd)):
        if lib.v2m_struct_size(which) != C.sizeof(cls):
            raise ImportError("ctypes layout of %s (%d B) differs from the C ABI (%d B)"
                              % (cls.__name__, C.sizeof(cls), lib.v2m_struct_size(which)))
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().v2m_last_error().decode("utf-8", "replace")
        raise RuntimeError("v2m_b200 kernel call failed (status %d): %s" % (rc, msg))


_device_checked = False


def require_device(t: torch.Tensor) -> None:
    """No CPU fallback: every compute op must be handed sm_100 CUDA tensors."""
    global _device_checked
    if not t.is_cuda:
        raise RuntimeError("video2music_b200 ops run only on a B200 (sm_100a) CUDA device; got a %s tensor. "
                           "There is no CPU fallback." % t.device)
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's stream: a tensor of another device would be an illegal address there
        raise RuntimeError("tensor lives on %s but the current CUDA device is cuda:%d; run under torch.cuda.device(...) "
                           "(one process per GPU)" % (t.device, torch.cuda.current_device()))
    if not _device_checked:
        with torch.cuda.device(t.device):
            if not load().v2m_device_ok():
                raise RuntimeError("video2music_b200 kernels are built for sm_100a only; current device is %s"
                                   % torch.cuda.get_device_name(t.device))
        _device_checked = True


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError("unsupported dtype %s (float32 / bfloat16 only)" % dt)


def count_launches(n: int) -> None:
    global _launches
    _launches += n


def launches() -> int:
    return _launches


def reset_launches() -> None:
    global _launches
    _launches = 0
