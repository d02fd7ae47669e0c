"""ctypes binding of libae_b200.so (the C ABI declared in include/ae_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AE_B200_LIB") or os.path.join(_HERE, "libae_b200.so")   # AE_B200_LIB: developer builds only

PREC_FP32, PREC_BF16 = 0, 1
BACKEND_TC, BACKEND_SIMT = 0, 1
OP_RAW, OP_BNRELU, OP_BNBWD, OP_SIGMOID_BWD, OP_SPLIT_BF16 = 0, 1, 2, 3, 4
EPI_STORE, EPI_BIAS_STATS, EPI_RELUBWD_STATS, EPI_BNRELU_SPLIT = 0, 1, 2, 3
PART_ENC, PART_DEC, PART_HEAD = 0, 1, 2
BNC_ROWS = 8
DP_UNIQUE_ID_BYTES = 128
DP_IPC_HANDLE_BYTES, DP_FLAG_BYTES = 64, 128

c_void_p, c_int, c_int64, c_float, c_size_t = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t


class Operand(C.Structure):
    _fields_ = [("src", c_void_p), ("src2", c_void_p), ("bnc", c_void_p), ("scalar", c_float), ("mode", c_int)]


class Epilogue(C.Structure):
    _fields_ = [("mode", c_int), ("bias", c_void_p), ("y", c_void_p), ("bnc", c_void_p), ("stats", c_void_p)]


class ConvGeom(C.Structure):
    _fields_ = [("batch", c_int), ("hs", c_int), ("ws", c_int), ("cb", c_int), ("cs", c_int)]


class EngineConfig(C.Structure):
    _fields_ = [("latent_dim", c_int), ("num_classes", c_int), ("max_batch", c_int), ("precision", c_int),
                ("backend", c_int)]


class AdamConfig(C.Structure):
    _fields_ = [("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("weight_decay", c_float)]


P = C.POINTER
_SIGS = {
    "ae_last_error": (C.c_char_p, []),
    "ae_abi_version": (c_int, []),
    "ae_device_supported": (c_int, [c_int]),
    "ae_packed_weight_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ae_pack_conv_weight": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ae_split_operand_bytes": (c_size_t, [c_int64, c_int]),
    "ae_split_operand": (c_int, [P(Operand), c_int, c_int64, c_void_p, c_int, c_void_p]),
    "ae_conv2d_s2_fwd": (c_int, [P(ConvGeom), P(Operand), c_void_p, P(Epilogue), c_void_p, c_int, c_int, c_void_p]),
    "ae_conv2d_s2_dgrad": (c_int, [P(ConvGeom), P(Operand), c_void_p, P(Epilogue), c_void_p, c_int, c_int, c_void_p]),
    "ae_conv2d_s2_wgrad_workspace_bytes": (c_size_t, [P(ConvGeom), c_int, c_int]),
    "ae_conv2d_s2_wgrad": (c_int, [P(ConvGeom), P(Operand), P(Operand), c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "ae_thin_gather_fwd": (c_int, [P(Operand), c_void_p, P(Epilogue), c_void_p, c_int, c_int, c_int, c_void_p]),
    "ae_thin_scatter_sigmoid_fwd": (c_int, [P(Operand), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                            c_void_p]),
    "ae_thin_wgrad": (c_int, [P(Operand), P(Operand), c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
    "ae_thin_wgrad_workspace_bytes": (c_size_t, [c_int]),
    "ae_thin_bwd_fused": (c_int, [P(Operand), P(Operand), c_void_p, P(Epilogue), c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_int, c_int, c_int, c_void_p]),
    "ae_bn_finalize": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ae_bn_bwd_reduce": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ae_linear_fwd": (c_int, [P(Operand), c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_size_t, c_void_p]),
    "ae_linear_bwd": (c_int, [P(Operand), c_int, c_void_p, c_void_p, c_void_p, P(Epilogue), c_int, c_void_p, c_void_p,
                              c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ae_linear_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ae_softmax_ce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ae_sigmoid_mse_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "ae_mlp_param_layout": (c_int64, [c_int, c_int, P(c_int64), P(c_int64)]),
    "ae_mlp_fwd_bwd_ce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_uint64, c_float,
                                  c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ae_mlp_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ae_mlp_train_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_uint64, c_void_p, c_float, c_int, c_int,
                                  c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, P(AdamConfig), c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "ae_mlp_train_step_indexed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          C.c_uint64, c_void_p, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_size_t, P(AdamConfig), c_void_p, c_void_p, c_void_p, c_void_p]),
    "ae_mlp_forward_eval": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ae_mlp_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ae_adam_step_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                  c_float, c_float, c_void_p, c_void_p]),
    "ae_augment_u8": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, C.c_uint64, c_float,
                              c_float, c_void_p, c_int, c_void_p]),
    "ae_layout_nchw_f32_to_nhwc_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ae_layout_nhwc_f32_to_nchw_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ae_layout_nchw_f32_to_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ae_layout_nhwc_bf16_to_nchw_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ae_engine_create": (c_int, [P(EngineConfig), P(c_void_p)]),
    "ae_engine_destroy": (None, [c_void_p]),
    "ae_engine_param_layout": (c_int, [c_void_p, c_int, P(c_int64), P(c_int64), P(c_int64)]),
    "ae_engine_bn_layout": (c_int, [c_void_p, c_int, P(c_int)]),
    "ae_engine_workspace_bytes": (c_size_t, [c_void_p]),
    "ae_engine_bind_workspace": (c_int, [c_void_p, c_void_p, c_size_t]),
    "ae_engine_bind_part": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ae_engine_pack_weights": (c_int, [c_void_p, c_int, c_void_p]),
    "ae_encoder_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ae_decoder_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ae_head_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ae_decoder_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ae_head_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ae_encoder_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "ae_train_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "ae_eval_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ae_step_graph_capture": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int64, P(AdamConfig), c_void_p, c_void_p, c_void_p, P(c_void_p)]),
    "ae_step_graph_launch": (c_int, [c_void_p, c_void_p]),
    "ae_step_graph_num_kernels": (c_int, [c_void_p]),
    "ae_step_graph_destroy": (None, [c_void_p]),
    "ae_dp_get_unique_id": (c_int, [c_void_p]),
    "ae_dp_init": (c_int, [c_void_p, c_int, c_int, P(c_void_p)]),
    "ae_dp_allreduce": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ae_dp_world": (c_int, [c_void_p]),
    "ae_dp_ipc_export": (c_int, [c_void_p, c_void_p, P(c_int64)]),
    "ae_dp_peers_attach": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    "ae_dp_max_ctas": (c_int, [c_void_p]),
    "ae_dp_destroy": (None, [c_void_p]),
}

_lib = None


def declared_symbols():
    return sorted(_SIGS)


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
            "ae_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ae_abi_version() != 1:
        raise RuntimeError("libae_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().ae_last_error()
        raise RuntimeError("libae_b200: " + (msg.decode() if msg else f"error {rc}"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
