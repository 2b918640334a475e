"""ctypes binding of libcnhead_sm100.so (C ABI declared in include/cnhead.h).

There is NO fallback: if the shared library is missing or an input is not a CUDA fp32
tensor, the call raises ``RuntimeError``.  PyTorch is used only for device memory and
streams; every kernel is launched by the library on ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNH_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libcnhead_sm100.so")

MAX_HEADS = 3
TOTALS = 24
SCALARS = 8
ANGLE_NONE, ANGLE_SIGMOID, ANGLE_PERIODIC, LIMB_SQRT, LIMB_L1 = 0, 1, 2, 3, 4
FLAG_ACCURATE_MATH, FLAG_NO_STASH, FLAG_DEFER_TOTALS = 1, 2, 4
SOFTMAX_ENTROPY, SOFTMAX_ENTROPY_ETA, SOFTMAX_MAX_SQUARE = 0, 1, 2


class Head(C.Structure):
    _fields_ = [("map", C.c_void_p), ("target", C.c_void_p), ("mask", C.c_void_p), ("grad", C.c_void_p),
                ("D", C.c_int32), ("angle_mode", C.c_int32), ("elementwise_mask", C.c_int32),
                ("weight", C.c_float), ("angle_weight", C.c_float), ("n_pairs", C.c_int32), ("pairs", C.c_void_p)]


class Cand(C.Structure):
    _fields_ = [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("K", C.c_int32), ("G", C.c_int32),
                ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class DetLossArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("M", C.c_int32),
                ("n_heads", C.c_int32), ("flags", C.c_int32), ("B_global", C.c_int32),
                ("hm_logits", C.c_void_p), ("hm_gt", C.c_void_p), ("prob", C.c_void_p), ("grad_hm", C.c_void_p),
                ("ind", C.c_void_p), ("hm_weight", C.c_float), ("_pad", C.c_int32),
                ("heads", Head * MAX_HEADS),
                ("scalars", C.c_void_p), ("totals", C.c_void_p), ("norm", C.c_void_p), ("norm_out", C.c_void_p),
                ("cand", C.POINTER(Cand))]


MAX_PEERS = 8
MAILBOX_BYTES = 8192
E_PEER = -6


class Peers(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("mailbox", C.c_void_p * MAX_PEERS),
                ("status", C.c_void_p), ("timeout_ms", C.c_uint32), ("_pad", C.c_uint32)]


class ScaleArgs(C.Structure):
    _fields_ = [("n_tensors", C.c_int32), ("_pad", C.c_int32), ("data", C.c_void_p * 4),
                ("count", C.c_int64 * 4), ("fa", C.c_void_p * 4), ("fb", C.c_void_p * 4)]


class DecodeArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("K", C.c_int32),
                ("D", C.c_int32), ("nk", C.c_int32), ("rotated", C.c_int32),
                ("heat", C.c_void_p), ("wh", C.c_void_p), ("reg", C.c_void_p), ("kps", C.c_void_p),
                ("dets", C.c_void_p), ("inds_out", C.c_void_p), ("kps_out", C.c_void_p),
                ("apply_sigmoid", C.c_int32), ("box_scale", C.c_float),
                ("counts_out", C.c_void_p), ("score_threshold", C.c_float), ("_pad", C.c_int32)]


class RasterArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("M", C.c_int32),
                ("min_overlap_num", C.c_int32), ("min_overlap_den", C.c_int32), ("_pad", C.c_int32),
                ("boxes", C.c_void_p), ("classes", C.c_void_p), ("n_obj", C.c_void_p), ("hm", C.c_void_p),
                ("wh", C.c_void_p), ("reg", C.c_void_p), ("ind", C.c_void_p), ("reg_mask", C.c_void_p)]


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"cnhead: CUDA extension not built: {LIB_PATH} is missing. Build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` or centernet-uda_b200/csrc/build.sh "
                f"(needs nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        L = C.CDLL(LIB_PATH)
        vp, st, sz, i32, i64, f32 = C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int64, C.c_float
        L.cnh_version.restype = C.c_int
        L.cnh_last_error.restype = C.c_char_p
        L.cnh_detloss_workspace_bytes.restype = sz
        L.cnh_detloss_workspace_bytes.argtypes = [C.POINTER(DetLossArgs)]
        for name in ("cnh_detloss_fused", "cnh_detloss_count", "cnh_detloss_main"):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(DetLossArgs), vp, sz, st]
        L.cnh_detloss_single_wave.restype = C.c_int
        L.cnh_detloss_single_wave.argtypes = [C.POINTER(DetLossArgs)]
        L.cnh_detloss_fused_peers.restype = C.c_int
        L.cnh_detloss_fused_peers.argtypes = [C.POINTER(DetLossArgs), C.POINTER(Peers), vp, sz, st]
        L.cnh_detloss_peers_finalize.restype = C.c_int
        L.cnh_detloss_peers_finalize.argtypes = [C.POINTER(DetLossArgs), C.POINTER(Peers), vp, sz, st]
        L.cnh_detloss_finalize.restype = C.c_int
        L.cnh_detloss_finalize.argtypes = [C.POINTER(DetLossArgs), vp, st]
        L.cnh_scale_inplace.restype = C.c_int
        L.cnh_scale_inplace.argtypes = [C.POINTER(ScaleArgs), st]
        L.cnh_softmax_workspace_bytes.restype = sz
        L.cnh_softmax_workspace_bytes.argtypes = [i32, i32, i32, i32]
        L.cnh_softmax_loss.restype = C.c_int
        L.cnh_softmax_loss.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, sz, st]
        L.cnh_entropy_map_fwd.restype = C.c_int
        L.cnh_entropy_map_fwd.argtypes = [vp, vp, i32, i32, i32, i32, st]
        L.cnh_entropy_map_bwd.restype = C.c_int
        L.cnh_entropy_map_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, st]
        L.cnh_bce_const.restype = C.c_int
        L.cnh_bce_const.argtypes = [vp, vp, vp, i64, f32, st]
        L.cnh_decode_workspace_bytes.restype = sz
        L.cnh_decode_workspace_bytes.argtypes = [C.POINTER(DecodeArgs)]
        L.cnh_decode.restype = C.c_int
        L.cnh_decode.argtypes = [C.POINTER(DecodeArgs), vp, sz, st]
        L.cnh_cand_workspace_bytes.restype = sz
        L.cnh_cand_workspace_bytes.argtypes = [i32]
        L.cnh_cand_state_bytes.restype = sz
        L.cnh_cand_state_bytes.argtypes = [C.POINTER(Cand)]
        L.cnh_decode_candidates.restype = C.c_int
        L.cnh_decode_candidates.argtypes = [C.POINTER(DecodeArgs), C.POINTER(Cand), st]
        L.cnh_copy_async.restype = C.c_int
        L.cnh_copy_async.argtypes = [vp, vp, sz, st]
        L.cnh_raster_targets.restype = C.c_int
        L.cnh_raster_targets.argtypes = [C.POINTER(RasterArgs), st]
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().cnh_last_error()
        raise RuntimeError(f"cnhead.{what} failed (code {rc}): {msg.decode() if msg else ''}")


try:                                                    # raw handle without building Stream objects (~10x cheaper)
    _raw_stream = torch._C._cuda_getCurrentRawStream
    _cur_device = torch._C._cuda_getDevice
except AttributeError:                                  # pragma: no cover -- older torch
    _raw_stream = _cur_device = None


def stream_ptr() -> int:
    """cudaStream_t of the calling thread's current stream on the current device."""
    if _raw_stream is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    """The product path has no CPU route: anything but a CUDA tensor of the right dtype is an error."""
    if not torch.is_tensor(t):
        raise RuntimeError(f"cnhead: {name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"cnhead: {name} is on {t.device}; this path runs on CUDA only (no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"cnhead: {name} must be {dtype}, got {t.dtype}")
    if _cur_device is not None and t.device.index != _cur_device():
        # the library launches on the CURRENT device's current stream: a tensor of another device would be touched
        # through a foreign pointer (illegal address, or silent peer access)
        raise RuntimeError(f"cnhead: {name} lives on cuda:{t.device.index} but the current device is cuda:{_cur_device()}; "
                           f"wrap the call in torch.cuda.device({t.device.index})")
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------
# persistent, self-cleaning workspaces: zero-filled once, kernels leave their counters zeroed.
# Keyed by (device, stream, kind) so that work on different streams never shares counters.
# --------------------------------------------------------------------------------------------
_workspaces: Dict[Tuple[int, int, str], torch.Tensor] = {}


def workspace(kind: str, nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(), kind)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        grow = max(int(nbytes), 4096)
        if ws is not None:
            grow = max(grow, 2 * ws.numel())
        ws = torch.zeros(grow, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def reset_workspaces() -> None:
    """Drop all cached workspaces (e.g. after a failed launch left counters dirty)."""
    _workspaces.clear()
