"""ctypes binding of libnmae.so (C ABI in include/nmae.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnmae.so")

# p = pointer, i = int, f = float, l = long long ; every function ends with (device:int, stream:void*)
_SIGS = {
    "nmae_pad_grid": "piiipii",
    "nmae_ingest_scene": "p" "iiiiiiii" "p" "ii",
    "nmae_patch_embed_fwd": "pppppppp" "iiii" "f" "ppppp",
    "nmae_patch_embed_bwd": "pppppppp" "iiii" "pppppp",
    "nmae_layernorm_fwd": "ppp" "ii" "f" "ppp",
    "nmae_layernorm_bwd": "ppppp" "ii" "pppp",
    "nmae_linear_fwd": "ppp" "iiii" "ppp" "i" "pp",
    "nmae_linear_bwd_input": "pp" "iiii" "ppp",
    "nmae_linear_bwd_weight": "pp" "iii" "pp",
    "nmae_linear_prep_batch": "p" "i" "l",
    "nmae_window_attention_fwd": "pp" "iiiiiii" "pp",
    "nmae_window_attention_bwd": "ppppp" "iiiiiii" "pp",
    "nmae_patch_merge_fwd": "pppp" "iiiii" "f" "ppppp",
    "nmae_patch_merge_bwd": "ppppppp" "iiiii" "pppppp",
    "nmae_convT_k_eq_s_fwd": "ppp" "iiiiiii" "p" "i" "p",
    "nmae_convT_k_eq_s_bwd": "p" "i" "pp" "iiiiiii" "pppp",
    "nmae_conv3_image_build": "p" "iiiiiiii" "p",
    "nmae_conv3_image_build_in_lrelu": "pp" "iiiii" "ff" "p",
    "nmae_conv3x3x3_fwd": "pppp" "iiiiii" "pp",
    "nmae_conv3x3x3_dgrad": "ppp" "iiiiii" "pp" "i",
    "nmae_conv3x3x3_wgrad": "pppp" "iiiiii" "ppp",
    "nmae_instnorm_stats": "p" "iii" "p",
    "nmae_in_lrelu_apply_fwd": "pppp" "iii" "ff" "p",
    "nmae_in_lrelu_apply_out_fwd": "pppp" "iii" "ff" "pppp",
    "nmae_in_lrelu_apply_bwd": "pppppp" "iii" "ff" "pppppp",
    "nmae_in_lrelu_apply_bwd_image": "pppppp" "iiiii" "ff" "pppppp",
    "nmae_conv3h_image_build": "p" "iiiiiii" "p" "ff" "pp",
    "nmae_conv3h_fwd": "ppp" "iiiiii" "pp",
    "nmae_conv3h_dgrad": "ppp" "iiiiii" "pp" "i",
    "nmae_conv3h_wgrad": "ppp" "iiiiii" "p",
    "nmae_in_lrelu_apply_bwd_image_h": "pppppp" "iiiii" "ff" "pppppppppppp",
    "nmae_copy_cols": "plpl" "l" "i",
    "nmae_upsample_nearest_add": "pp" "iiiiiiii",
    "nmae_upsample_trilinear_fwd": "pp" "iiiiiiii",
    "nmae_upsample_trilinear_bwd": "pp" "iiiiiiii",
    "nmae_colsum": "p" "l" "i" "l" "p",
    "nmae_scale_rows": "ppp" "i" "l" "i",
    "nmae_mae_loss_fwd": "pppp" "iii" "pp",
    "nmae_mae_loss_bwd": "pppp" "iii" "ppp",
    "nmae_multi_sumsq": "p" "i" "p",
    "nmae_multi_copy": "p" "i",
    "nmae_adamw_clip_step": "p" "i" "p" "fffffffff",
}
# workspace-size queries: name -> number of int arguments (all return long long bytes)
_WS_QUERIES = {"nmae_linear_weight_ws_bytes": 2, "nmae_patch_embed_weight_ws_bytes": 2, "nmae_patch_embed_bwd_ws_bytes": 4,
               "nmae_patch_merge_weight_ws_bytes": 1, "nmae_patch_merge_bwd_ws_bytes": 5, "nmae_convT_weight_ws_bytes": 3,
               "nmae_conv3x3x3_weight_ws_bytes": 2, "nmae_window_attention_lse_bytes": 5, "nmae_instnorm_stats_bytes": 2,
               "nmae_in_lrelu_bwd_sums_ws_bytes": 2}
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "f": ctypes.c_float, "l": ctypes.c_longlong}

_lib = None
launches = 0  # number of C-ABI calls issued (each enqueues >= 1 kernel); bench.py reports kernel counts separately


def exported_symbols():
    return ["nmae_version", "nmae_last_error", "nmae_launch_count", "nmae_window_attention_num_windows",
            "nmae_conv3_image_bytes", "nmae_conv3h_image_bytes", "nmae_conv3h_weight_ws_bytes", "nmae_linear_blob_layout"] + \
        list(_WS_QUERIES) + list(_SIGS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "nerf-mae_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.nmae_last_error.restype = ctypes.c_char_p
        L.nmae_version.restype = ctypes.c_int
        L.nmae_launch_count.restype = ctypes.c_ulonglong
        L.nmae_conv3_image_bytes.restype = ctypes.c_longlong
        L.nmae_conv3_image_bytes.argtypes = [ctypes.c_int] * 5
        L.nmae_window_attention_num_windows.argtypes = [ctypes.c_int] * 3
        L.nmae_conv3h_image_bytes.restype = ctypes.c_longlong
        L.nmae_conv3h_image_bytes.argtypes = [ctypes.c_int] * 5
        L.nmae_conv3h_weight_ws_bytes.restype = ctypes.c_longlong
        L.nmae_conv3h_weight_ws_bytes.argtypes = [ctypes.c_int] * 2
        L.nmae_linear_blob_layout.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)] * 2 + [ctypes.c_int]
        L.nmae_linear_blob_layout.restype = ctypes.c_int
        for name, nargs in _WS_QUERIES.items():
            fn = getattr(L, name)
            fn.argtypes = [ctypes.c_int] * nargs
            fn.restype = ctypes.c_longlong
        for name, sig in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = [_CT[c] for c in sig] + [ctypes.c_int, ctypes.c_void_p]
            fn.restype = ctypes.c_int
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a


def call(name: str, *args, device: torch.device):
    """Invoke a C-ABI entry point on torch's current stream of `device`."""
    global launches
    if device.type != "cuda":
        raise RuntimeError(f"{name}: nerf-mae_b200 runs on CUDA (sm_100a) only, got a tensor on {device}")
    L = lib()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(idx).cuda_stream
    if timed_calls is not None and name in timed_calls:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(idx))
        rc = getattr(L, name)(*[_ptr(a) for a in args], idx, stream)
        e1.record(torch.cuda.current_stream(idx))
        timed_calls[name].append((e0, e1, tuple(a for a in args if isinstance(a, int))))
    else:
        rc = getattr(L, name)(*[_ptr(a) for a in args], idx, stream)
    launches += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {L.nmae_last_error().decode()}")


def kernel_launches() -> int:
    """CUDA kernels launched by libnmae.so so far in this process."""
    return int(lib().nmae_launch_count())


# optional per-call CUDA-event timing (bench.py): {name: [(start_event, end_event, args), ...]}
timed_calls = None


def conv3_image_bytes(B: int, X: int, Y: int, Z: int, C: int) -> int:
    """Size of the tensor-core operand image of a (B,X,Y,Z,C) volume; 0 when C is not a multiple of 48."""
    return int(lib().nmae_conv3_image_bytes(B, X, Y, Z, C))


def num_windows(H: int, W: int, D: int) -> int:
    return lib().nmae_window_attention_num_windows(H, W, D)


def conv3h_image_bytes(B: int, X: int, Y: int, Z: int, C: int) -> int:
    """Size of the fp16 operand image of a (B,X,Y,Z,C) volume; 0 when C is a multiple of neither 48 nor 64."""
    return int(lib().nmae_conv3h_image_bytes(B, X, Y, Z, C))


def workspace_bytes(name: str, *dims: int) -> int:
    """Bytes of a caller-provided scratch buffer: workspace_bytes("nmae_linear_weight_ws_bytes", N, K) etc. (include/nmae.h)."""
    if name not in _WS_QUERIES:
        raise KeyError(name)
    return int(getattr(lib(), name)(*[int(d) for d in dims]))


def linear_blob_layout(M: int, N: int, K: int, device_index: int):
    """(tile_n, k_group) of the tensor-core GEMM out[M,N] = A[M,K] W^T, or (0, 0) when it takes the CUDA-core path."""
    nt, kg = ctypes.c_int(0), ctypes.c_int(0)
    rc = lib().nmae_linear_blob_layout(int(M), int(N), int(K), ctypes.byref(nt), ctypes.byref(kg), int(device_index))
    if rc != 0:
        raise RuntimeError(f"nmae_linear_blob_layout failed ({rc}): {lib().nmae_last_error().decode()}")
    return nt.value, kg.value
