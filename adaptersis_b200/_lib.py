"""ctypes binding of libasis_b200.so -- the thin C-ABI "torch extension".

PyTorch is plumbing here: it owns device memory and streams; every kernel on the hot path is a
hand-written sm_100a kernel reached through the C ABI declared in include/asis_b200.h.  There is
no CPU fallback and no alternative backend: if the library is missing, importing any op raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ASIS_LIB") or os.path.join(_HERE, "libasis_b200.so")   # ASIS_LIB: a trace build (tools/)

F32, BF16 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
EPI_NONE, EPI_GELU, EPI_SCALE_RESIDUAL, EPI_DGELU, EPI_ACCUMULATE, EPI_GELU_GRAD, EPI_MUL_AUX = 0, 1, 2, 3, 4, 5, 6

_DT = {torch.float32: F32, torch.bfloat16: BF16}

c_void_p, c_int, c_i64, c_size_t, c_float = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t,
                                             ctypes.c_float)

# name -> (restype, argtypes); mirrors include/asis_b200.h one to one
_P, _I, _L, _Z, _F = c_void_p, c_int, c_i64, c_size_t, c_float
SIGNATURES = {
    "asis_abi_version": (_I, []),
    "asis_last_error": (ctypes.c_char_p, []),
    "asis_launch_count": (ctypes.c_uint64, []),
    "asis_msda_forward": (_I, [_P, _I, _P, _P, _P, _P, _P, _I] + [_I] * 7 + [_P]),
    "asis_msda_backward_workspace_bytes": (_Z, [_I] * 7),
    "asis_msda_backward": (_I, [_P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P] + [_I] * 7 + [_P, _Z, _P]),
    "asis_msda_prep_forward": (_I, [_P, _P, _I, _P, _I, _I, _I, _P, _P, _P] + [_I] * 5 + [_P]),
    "asis_msda_prep_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _I] + [_I] * 5 + [_P]),
    "asis_layernorm_forward": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _F, _P]),
    "asis_layernorm_backward_workspace_bytes": (_Z, [_I, _I]),
    "asis_layernorm_backward": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "asis_gemm": (_I, [_I, _P, _I, _L, _P, _I, _L, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _P, _I, _L, _P]),
    "asis_colsum_workspace_bytes": (_Z, [_I, _I]),
    "asis_colsum": (_I, [_P, _I, _P, _I, _L, _P, _I, _I, _I, _P, _Z, _P]),
    "asis_layerscale_backward_workspace_bytes": (_Z, [_I, _I]),
    "asis_layerscale_backward": (_I, [_P, _P, _P, _P, _I, _P, _P, _I, _I, _P, _Z, _P]),
    "asis_scale_cols": (_I, [_P, _I, _P, _P, _I, _L, _I, _P]),
    "asis_add": (_I, [_P, _I, _P, _I, _P, _I, _L, _P]),
    "asis_cast": (_I, [_P, _I, _P, _I, _L, _P]),
    "asis_attention_forward_workspace_bytes": (_Z, [_I] * 5),
    "asis_attention_forward": (_I, [_I, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "asis_attention_backward_workspace_bytes": (_Z, [_I] * 5),
    "asis_attention_backward": (_I, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "asis_patchify": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _L, _P]),
    "asis_frames_to_batch": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "asis_sgd_step": (_I, [_I, _P, _P, _P, _P, _F, _F, _F, _P]),
    "asis_upsample2x_bilinear_forward": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "asis_upsample2x_bilinear_backward": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "asis_upsample2x_bilinear_forward_padded": (_I, [_P, _P] + [_I] * 7 + [_P]),
    "asis_upsample2x_bilinear_backward_padded": (_I, [_P, _P] + [_I] * 7 + [_P]),
    "asis_im2col": (_I, [_P, _I, _P, _I] + [_I] * 8 + [_L, _P]),
    "asis_col2im": (_I, [_P, _P, _I] + [_I] * 8 + [_L, _P]),
    "asis_conv3x3s1_gemm": (_I, [_I, _P, _P, _P, _I, _P] + [_I] * 5 + [_P]),
    "asis_chan_stats_workspace_bytes": (_Z, [_I] * 4),
    "asis_chan_stats": (_I, [_I, _P, _P, _I] + [_I] * 6 + [_P] * 5 + [_I, _P, _P, _P, _Z, _P]),
    "asis_bn_apply": (_I, [_I, _P, _P, _I, _P, _I] + [_I] * 7 + [_P] * 6 + [_I, _P]),
    "asis_maxpool3x3s2_forward": (_I, [_P, _P, _P, _I] + [_I] * 6 + [_P]),
    "asis_maxpool3x3s2_backward": (_I, [_P, _P, _P, _I] + [_I] * 6 + [_P]),
    "asis_seg_head_workspace_bytes": (_Z, [_I] * 5),
    "asis_seg_head_forward": (_I, [_P, _I, _P, _P, _P] + [_I] * 5 + [_P, _Z, _P]),
    "asis_seg_head_backward": (_I, [_P, _I, _P, _P, _P, _P] + [_I] * 5 + [_P, _Z, _P]),
    "asis_dwconv3x3_forward": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P]),
    "asis_dwconv3x3_backward_workspace_bytes": (_Z, [_I, _I, _I]),
    "asis_dwconv3x3_backward": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _Z, _P]),
}

_lib = None


def load():
    """Load the shared library (once).  Fails loudly: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m adaptersis_b200.build` or `__graft_entry__.build()`); "
            "adaptersis_b200 has no CPU or library fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.asis_abi_version() != 1:
        raise RuntimeError("libasis_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().asis_last_error().decode(errors="replace")
        if rc == -1 and ("must be" in msg and "reference_points" in msg):
            raise ValueError(msg)
        raise RuntimeError(f"libasis_b200 error {rc}: {msg}")


def dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype} (float32 / bfloat16 only)") from None


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("adaptersis_b200 ops run on CUDA tensors only (no CPU fallback)")


_ws = {}


def workspace(nbytes, device):
    """Per-(device, stream) scratch buffer, grown on demand.  Stream-ordered reuse is safe
    because every consumer is enqueued on the same stream."""
    if nbytes == 0:
        return None
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def launch_count():
    return int(load().asis_launch_count())
