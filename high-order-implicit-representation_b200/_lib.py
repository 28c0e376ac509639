"""ctypes binding of libhoir_b200.so (the C ABI declared in include/hoir.h).

There is no CPU fallback: every compute entry point takes raw CUDA device pointers, and the
loader raises if the library is missing on a machine that has a GPU (or whenever a compute
call is attempted without it).  The oracle under ``oracle/`` is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# HOIR_LIB: load another build of the same sources instead (libhoir_b200_dbg.so = -DHOIR_DEBUG_BOUNDS, see csrc/Makefile)
LIB_PATH = os.environ.get("HOIR_LIB") or os.path.join(_PKG_DIR, "libhoir_b200.so")
DEBUG_LIB_PATH = os.path.join(_PKG_DIR, "libhoir_b200_dbg.so")
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

HOIR_MAX_LAYERS = 8
HOIR_MAX_N = 32
HOIR_MAX_ROTATIONS = 8

HOIR_OK, HOIR_EINVAL, HOIR_EUNSUPPORTED, HOIR_ECUDA = 0, 1, 2, 3
KIND = {"continuous": 0, "discontinuous": 1, "polynomial": 2, "fourier": 3}
TARGET_F32, TARGET_U8_IMAGE = 0, 1


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("segments", C.c_int),
                ("in_features", C.c_int), ("out_features", C.c_int),
                ("length", C.c_float), ("rescale", C.c_float), ("periodicity", C.c_float),
                ("fourier_arg_scale", C.c_float), ("fourier_odd_is_sin", C.c_int)]


class MlpDesc(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("layers", LayerDesc * HOIR_MAX_LAYERS),
                ("normalize", C.c_int), ("norm_eps", C.c_float)]


class MeshDesc(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("rotations", C.c_int),
                ("normalize_mode", C.c_int), ("first_pixel", C.c_longlong),
                ("rot_x", C.c_float * HOIR_MAX_ROTATIONS), ("rot_y", C.c_float * HOIR_MAX_ROTATIONS),
                ("rot_sum", C.c_float * HOIR_MAX_ROTATIONS)]


OPT_KIND = {"adam": 1, "sparse_lion": 2}


class OptDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("weight_decay", C.c_float), ("step", C.c_longlong),
                ("flat_w", C.c_void_p), ("flat_g", C.c_void_p), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("packed", C.c_void_p), ("loss_out", C.c_void_p),
                ("keep_grad", C.c_int)]


HOIR_MAX_PEERS = 16


class PeerDesc(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("block", C.c_void_p * HOIR_MAX_PEERS)]


class HoirUnsupported(RuntimeError):
    """The fused whole-network kernel has no instantiation for this network."""


_EXPORTS = {
    "hoir_last_error": (C.c_char_p, []),
    "hoir_version": (C.c_int, []),
    "hoir_shutdown": (C.c_int, []),
    "hoir_layer_nodes": (C.c_int, [C.POINTER(LayerDesc)]),
    "hoir_layer_forward": (C.c_int, [C.POINTER(LayerDesc), C.c_longlong, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "hoir_layer_backward": (C.c_int, [C.POINTER(LayerDesc), C.c_longlong, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hoir_segment_index": (C.c_int, [C.POINTER(LayerDesc), C.c_longlong, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "hoir_maxabs_forward": (C.c_int, [C.c_longlong, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "hoir_maxabs_backward": (C.c_int, [C.c_longlong, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "hoir_mesh_positions": (C.c_int, [C.POINTER(MeshDesc), C.c_longlong, C.c_void_p, C.c_void_p]),
    "hoir_neighborhood_patches": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hoir_neighborhood_gather": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "hoir_neighborhood_fold": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "hoir_radial_uniforms": (C.c_int, [C.c_longlong, C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "hoir_radial_indices": (C.c_int, [C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.c_void_p]),
    "hoir_radial_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "hoir_mlp_fused_supported": (C.c_int, [C.POINTER(MlpDesc)]),
    "hoir_mlp_forward_supported": (C.c_int, [C.POINTER(MlpDesc)]),
    "hoir_mlp_forward_acts": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "hoir_interp_frame": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hoir_mlp_packed_size": (C.c_longlong, [C.POINTER(MlpDesc)]),
    "hoir_mlp_packed_offset": (C.c_longlong, [C.POINTER(MlpDesc), C.c_int]),
    "hoir_mlp_pack": (C.c_int, [C.POINTER(MlpDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "hoir_mlp_forward": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p,
                                   C.POINTER(MeshDesc), C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p]),
    "hoir_mlp_train_step": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p,
                                      C.POINTER(MeshDesc), C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_float,
                                      C.c_void_p]),
    "hoir_mlp_optimizer_tail": (C.c_int, [C.POINTER(MlpDesc), C.POINTER(OptDesc), C.c_void_p]),
    "hoir_mlp_train_step_fused": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p,
                                            C.POINTER(MeshDesc), C.c_int, C.c_void_p, C.c_float,
                                            C.POINTER(OptDesc), C.c_void_p]),
    "hoir_mlp_order_supported": (C.c_int, [C.POINTER(MlpDesc)]),
    "hoir_mlp_batch_sort": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p, C.POINTER(MeshDesc), C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_float), C.c_void_p]),
    "hoir_mlp_train_step_order": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p, C.POINTER(MeshDesc),
                                            C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_void_p), C.c_void_p, C.c_float, C.c_void_p]),
    "hoir_mlp_train_step_fused_order": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p, C.POINTER(MeshDesc),
                                                  C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.POINTER(OptDesc),
                                                  C.POINTER(PeerDesc), C.c_void_p]),
    "hoir_peer_stride": (C.c_longlong, [C.POINTER(MlpDesc)]),
    "hoir_peer_block_floats": (C.c_longlong, [C.POINTER(MlpDesc)]),
    "hoir_mlp_train_step_fused_peer": (C.c_int, [C.POINTER(MlpDesc), C.c_longlong, C.c_void_p,
                                                 C.POINTER(MeshDesc), C.c_int, C.c_void_p, C.c_float,
                                                 C.POINTER(OptDesc), C.POINTER(PeerDesc), C.c_void_p]),
    "hoir_adam_step": (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_longlong, C.c_void_p]),
    "hoir_sparse_lion_step": (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "hoir_adam_step_dev": (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "hoir_sparse_lion_step_dev": (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "hoir_write_floats": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "hoir_tf32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_void_p]),
    "hoir_ffma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                 C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_EXPORTS)

_lib: Optional[C.CDLL] = None

#: bumped whenever a kernel of this library writes weights through raw pointers (optimizer steps, CUDA-graph replays):
#: tensor version counters do not see those writes, caches of kernel-side weight copies key on this as well
WEIGHT_EPOCH = 0


def weights_changed() -> None:
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libhoir_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    subprocess.run(["make", "-C", CSRC_DIR, "-j8"], check=True, capture_output=not verbose)
    return os.path.join(_PKG_DIR, "libhoir_b200.so")


def build_debug(verbose: bool = False) -> str:
    """The same sources with -DHOIR_DEBUG_BOUNDS (asserts on the swizzled smem / TMEM addressing) -> libhoir_b200_dbg.so."""
    subprocess.run(["make", "-C", CSRC_DIR, "-j8", "debug"], check=True, capture_output=not verbose)
    return DEBUG_LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  hoir_b200 has no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code == HOIR_OK:
        return
    msg = lib().hoir_last_error().decode("utf-8", "replace")
    if code == HOIR_EINVAL:
        raise ValueError(msg)
    if code == HOIR_EUNSUPPORTED:
        raise HoirUnsupported(msg)
    raise RuntimeError(msg)


def stream_ptr(device: Optional[torch.device] = None) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dptr(t: Optional[torch.Tensor], dtype=torch.float32) -> C.c_void_p:
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("hoir_b200 has no CPU path: expected a CUDA tensor, got a "
                           f"{t.device.type} tensor")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def ptr_table(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * HOIR_MAX_LAYERS)()
    for k, t in enumerate(tensors):
        arr[k] = dptr(t).value
    return arr


def make_layer_desc(kind: str, n: int, segments: int, in_features: int, out_features: int,
                    length: float = 2.0, rescale_output: bool = False,
                    periodicity: Optional[float] = None) -> LayerDesc:
    from . import spec
    if kind not in KIND:
        raise ValueError(f"Fully connected layer type {kind} not recognized. "
                         f"Must be one of {list(KIND)}")
    return LayerDesc(KIND[kind], int(n), int(segments or 1), int(in_features), int(out_features),
                     float(length), (1.0 / in_features) if rescale_output else 1.0,
                     float(periodicity) if periodicity else 0.0,
                     float(spec.FOURIER_ARG_SCALE), int(spec.FOURIER_ODD_IS_SIN))


def make_mlp_desc(layers: List[LayerDesc], normalize: bool, eps: float) -> MlpDesc:
    if len(layers) > HOIR_MAX_LAYERS:
        raise ValueError(f"at most {HOIR_MAX_LAYERS} layers")
    m = MlpDesc()
    m.num_layers = len(layers)
    for k, d in enumerate(layers):
        m.layers[k] = d
    m.normalize = 1 if normalize else 0
    m.norm_eps = eps
    return m


def make_mesh_desc(rows: int, cols: int, rotations: int, normalize: bool = True,
                   first_pixel: int = 0) -> MeshDesc:
    import math
    from . import spec
    if rotations < 1 or rotations > HOIR_MAX_ROTATIONS:
        raise ValueError(f"rotations must be in [1, {HOIR_MAX_ROTATIONS}]")
    m = MeshDesc()
    m.rows, m.cols, m.rotations = rows, cols, rotations
    m.normalize_mode = 0 if not normalize else (1 if spec.MESH_NORMALIZE == "max_size" else 2)
    m.first_pixel = first_pixel
    for i in range(rotations):
        theta = (math.pi / 2.0) * (i / rotations)
        rx, ry = math.cos(theta), math.sin(theta)
        m.rot_x[i], m.rot_y[i], m.rot_sum[i] = rx, ry, math.fabs(rx) + math.fabs(ry)
    return m
