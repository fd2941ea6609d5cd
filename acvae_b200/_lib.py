"""ctypes binding of the C-ABI shared library `libacvae_b200.so`
(`include/acvae_b200.h`).  There is NO fallback: if the library is missing or
was built for a different ABI version, importing the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacvae_b200.so")
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 2

c_float_p = C.POINTER(C.c_float)
c_void_p = C.c_void_p


class Dims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ("N", "Te", "T", "E", "A", "V", "Eenc", "L", "mem_rep", "variant")]


_W_FIELDS = [
    ("ln_w", 1), ("ln_b", 1), ("q_emb", 1), ("q_wih", 2), ("q_whh", 2), ("q_bih", 2), ("q_bhh", 2),
    ("q_head_w", 1), ("q_head_b", 1), ("p_emb", 1), ("p_attn_w", 1), ("p_attn_b", 1), ("p_attn_v", 1),
    ("p_wih", 1), ("p_whh", 1), ("p_bih", 1), ("p_bhh", 1), ("p_head_w", 1), ("p_head_b", 1),
    ("d_emb", 1), ("d_attn_w", 1), ("d_attn_b", 1), ("d_attn_v", 1), ("d_wih", 1), ("d_whh", 1),
    ("d_bih", 1), ("d_bhh", 1), ("cls_w", 1), ("cls_b", 1), ("g_w", 1), ("g_b", 1),
]


class Weights(C.Structure):
    _fields_ = [(k, c_void_p if n == 1 else c_void_p * n) for k, n in _W_FIELDS]


class WeightGrads(C.Structure):
    _fields_ = [(k, c_void_p if n == 1 else c_void_p * n) for k, n in _W_FIELDS]


class TrainIO(C.Structure):
    _fields_ = [(k, c_void_p) for k in (
        "audio_embeds", "mem_lens", "caps_ids", "cap_lens", "eps_q", "eps_p", "tf_flags", "dis_flags",
        "q_means", "q_logs", "q_z", "q_means_utt", "p_means", "p_logs", "p_z", "outputs", "p_means_utt",
        "attn_weights", "rnn_input", "seqs", "sampled_logprobs", "logit_lse", "logit_sum", "logits")]


class TrainGradsIn(C.Structure):
    _fields_ = [(k, c_void_p) for k in (
        "d_q_means", "d_q_logs", "d_q_z", "d_p_means", "d_p_logs", "d_p_z", "d_outputs",
        "d_q_means_utt", "d_p_means_utt")]


class SampleIO(C.Structure):
    _fields_ = [("audio_embeds", c_void_p), ("mem_lens", c_void_p), ("eps_p", c_void_p), ("u", c_void_p),
                ("method", C.c_int32), ("temp", C.c_float), ("start_idx", C.c_int32), ("end_idx", C.c_int32),
                ("seqs", c_void_p), ("sampled_logprobs", c_void_p), ("p_means", c_void_p), ("p_logs", c_void_p),
                ("p_z", c_void_p), ("outputs", c_void_p), ("n_steps", c_void_p), ("rng_state", c_void_p)]


# every symbol include/acvae_b200.h declares: (name, restype, argtypes)
_i32, _i64, _f, _sz, _vp = C.c_int32, C.c_int64, C.c_float, C.c_size_t, c_void_p
_DP, _WP = C.POINTER(Dims), C.POINTER(Weights)
SYMBOLS = {
    "acvae_last_error": (C.c_char_p, []),
    "acvae_abi_version": (C.c_int, []),
    "acvae_launch_count": (C.c_uint64, []),
    "acvae_set_kernel_probe": (C.c_int, [C.c_char_p, _vp, _vp]),
    "acvae_kernel_probe_hits": (C.c_int, []),
    "acvae_gemm": (C.c_int, [_i32, _i32, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _i64, _i32,
                             C.POINTER(C.c_int32), _vp]),
    "acvae_train_workspace_bytes": (_sz, [_DP]),
    "acvae_memory_prepare": (C.c_int, [_DP, _WP, _vp, _vp, _vp, _vp, _vp]),
    "acvae_train_fwd": (C.c_int, [_DP, _WP, C.POINTER(TrainIO), _vp, _sz, _vp]),
    "acvae_train_bwd": (C.c_int, [_DP, _WP, C.POINTER(TrainIO), C.POINTER(TrainGradsIn),
                                  C.POINTER(WeightGrads), _vp, _vp, _sz, _vp]),
    "acvae_defer_classifier_grads": (C.c_int, [_i32]),
    "acvae_join_deferred": (C.c_int, [_vp]),
    "acvae_vocab_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "acvae_vocab_logits": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "acvae_vocab_logits_bwd": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "acvae_vocab_stats": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "acvae_vocab_ce_fwd": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _f, _i32, _vp, _vp, _vp,
                                     _vp, _sz, _vp]),
    "acvae_vocab_ce_bwd": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _sz, _vp]),
    "acvae_kl_fwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "acvae_kl_bwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "acvae_sample_workspace_bytes": (_sz, [_DP]),
    "acvae_decode_sample": (C.c_int, [_DP, _WP, C.POINTER(SampleIO), _vp, _sz, _vp]),
    "acvae_beam_workspace_bytes": (_sz, [_DP, _i32]),
    "acvae_beam_search": (C.c_int, [_DP, _WP, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _sz, _vp]),
    "acvae_dbs_workspace_bytes": (_sz, [_DP, _i32, _i32]),
    "acvae_diverse_beam_search": (C.c_int, [_DP, _WP, _vp, _vp, _vp, _i32, _i32, _f, _f, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "acvae_loss_combine_fwd": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp]),
    "acvae_loss_combine_bwd": (C.c_int, [_i64, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "acvae_diversity_stats": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "acvae_set_input_event": (C.c_int, [_vp]),
    "acvae_ipc_export": (C.c_int, [_vp, _vp, C.POINTER(C.c_int64)]),
    "acvae_ipc_open": (C.c_int, [_vp, _i64, C.POINTER(_vp)]),
    "acvae_ipc_close_all": (C.c_int, []),
    "acvae_dp_comm_bytes": (_sz, []),
    "acvae_dp_workspace_bytes": (_sz, []),
    "acvae_dp_clip_adam": (C.c_int, [_i32, _i32, _i64, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _sz, _vp]),
    "acvae_set_bucket_event": (C.c_int, [_vp]),
    "acvae_mbleu_stats": (C.c_int, [_i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "acvae_encoder_handoff_fwd": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "acvae_encoder_handoff_bwd": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "acvae_debug_set_chain_trace": (C.c_int, [_vp]),
    "acvae_set_precision": (C.c_int, [_i32]),
    "acvae_get_precision": (C.c_int, []),
    "acvae_clip_adam_workspace_bytes": (_sz, []),
    "acvae_clip_adam": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _vp, _vp, _i32, _vp, _sz, _vp]),
    "acvae_clip_adam_dev": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build(verbose: bool = False) -> str:
    """Compile csrc/capi.cu for sm_100a into the in-tree shared library."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs) \
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(os.path.join(_HERE, "..", "include", "acvae_b200.h")):
        return LIB_PATH
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH, os.path.join(CSRC, "capi.cu")]
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library.  Raises if it is absent: the product has no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(acvae_b200 has no CPU or eager fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.acvae_abi_version() != ABI_VERSION:
            raise RuntimeError("libacvae_b200.so ABI version mismatch; rebuild")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"acvae_b200 {what} failed: {lib().acvae_last_error().decode()}")
