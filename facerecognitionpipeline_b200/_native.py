"""ctypes binding of libfrb200.so (C ABI declared in include/frb200.h).

There is deliberately no fallback: if the shared library is missing, or the device is not an
sm_100 GPU, every entry point raises.  The product path never routes through `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libfrb200.so"
CSRC_DIR = _PKG_DIR / "csrc"

FRB_OP_STEM, FRB_OP_CONV, FRB_OP_FC = 0, 1, 2
FRB_EMBED_L2, FRB_EMBED_RENORM, FRB_EMBED_FLIP = 1, 2, 4


class LayerDesc(C.Structure):
    """Mirror of `frb_layer_desc` (include/frb200.h)."""

    _fields_ = [
        ("op", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("hin", C.c_int32), ("win", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("in_buf", C.c_int32), ("out_buf", C.c_int32),
        ("sc_buf", C.c_int32), ("sc_cin", C.c_int32), ("sc_hin", C.c_int32), ("sc_win", C.c_int32),
        ("sc_stride", C.c_int32),
        ("res_buf", C.c_int32), ("res_h", C.c_int32), ("res_w", C.c_int32), ("res_stride", C.c_int32),
        ("bias_cases", C.c_int32),
        ("has_prelu", C.c_int32),
        ("reserved", C.c_int32 * 2),
        ("w_off", C.c_int64), ("w_bytes", C.c_int64),
        ("bias_off", C.c_int64),
        ("prelu_off", C.c_int64),
    ]


class TrackResult(C.Structure):
    """Mirror of `frb_track_result`."""

    _fields_ = [
        ("winner", C.c_int64), ("confidence", C.c_double), ("consensus_strength", C.c_double),
        ("num_quality_frames", C.c_int32), ("total_frames_evaluated", C.c_int32),
        ("candidate", C.c_int64), ("candidate_confidence", C.c_double),
        ("candidate_num_quality_frames", C.c_int32), ("recognized", C.c_int32),
    ]


class WarpJob(C.Structure):
    """Mirror of `frb_warp_job`."""

    _fields_ = [
        ("src_off", C.c_uint64),
        ("H", C.c_int32), ("W", C.c_int32), ("pitch", C.c_int32), ("_pad", C.c_int32),
        ("M", C.c_double * 6),
    ]


_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes); also the list the CPU-only test checks for exported symbols
SIGNATURES = {
    "frb_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "frb_ctx_destroy": (None, [_vp]),
    "frb_last_error": (C.c_char_p, [_vp]),
    "frb_launch_count": (_ll, [_vp]),
    "frb_preprocess_u8": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp]),
    "frb_warp_normalize": (_i, [_vp, _vp, C.POINTER(WarpJob), _i, _i, _vp, _vp, _vp]),
    "frb_backbone_load": (_i, [_vp, C.POINTER(LayerDesc), _i, _vp, C.c_size_t, _i]),
    "frb_embed": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "frb_backbone_flops_per_face": (C.c_double, [_vp]),
    "frb_embed_schedule": (_i, [_vp, _vp]),
    "frb_gallery_upload": (_i, [_vp, _vp, _ll, _ll, _i]),
    "frb_gallery_size": (_ll, [_vp]),
    "frb_gallery_generation": (_ll, [_vp]),
    "frb_backbone_generation": (_ll, [_vp]),
    "frb_gallery_upload_samples": (_i, [_vp, _vp, _ll, _vp, _ll, _i]),
    "frb_identity_scores": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _ll, _vp]),
    "frb_match_identities": (_i, [_vp, _vp, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "frb_aggregate_templates": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    "frb_track_consensus": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, C.c_double, _i, C.c_double, _vp, _vp]),
    "frb_best_frames": (_i, [_vp, _vp, _vp, _vp, _i, C.c_double, _vp, _vp, _vp, _vp]),
    "frb_match": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "frb_match_profile": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "frb_match_last_flagged": (_i, [_vp]),
    "frb_topk_merge": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "frb_topk_merge_packed": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "frb_match_packed": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp]),
    "frb_xchg_create": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "frb_xchg_connect": (_i, [_vp, _vp]),
    "frb_xchg_status": (_i, [_vp]),
    "frb_xchg_local_buffer": (_vp, [_vp]),
    "frb_xchg_connect_local": (_i, [_vp, _i, _vp]),
    "frb_match_sharded": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "frb_prefetch_host": (_i, [_vp, _vp, _i, _i]),
    "frb_embed_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "frb_match_host": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp]),
    "frb_embed_match_host": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "frb_embed_profile": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "frb_debug_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "frb_debug_conv": (_i, [_vp, C.POINTER(LayerDesc), _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "frb_debug_stream_bw": (_i, [_vp, _vp, C.c_size_t, _i, _i, _vp]),
    "frb_debug_shift_mma": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "frb_debug_mma_rate": (_i, [_vp, _i, _i, _i, _vp]),
    "frb_debug_mma2_rate": (_i, [_vp, _i, _i, _i, _vp]),
    "frb_debug_im2col": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def build(force: bool = False) -> Path:
    """Compile libfrb200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force and LIB_PATH.exists():
        LIB_PATH.unlink()
    proc = subprocess.run(["make", "-C", str(CSRC_DIR)], capture_output=True, text=True)
    if proc.returncode != 0 or not LIB_PATH.exists():
        raise NativeError(f"building libfrb200.so failed:\n{proc.stdout}\n{proc.stderr}")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (never builds implicitly, never falls back)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C facerecognitionpipeline_b200/csrc`). There is no CPU fallback.")
        # FRB_LIBRARY: load another build of the SAME library (A/B runs of compile-time variants); never a fallback
        l = C.CDLL(os.environ.get("FRB_LIBRARY") or str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class Context:
    """Owns one `frb_ctx` (one per process / GPU)."""

    def __init__(self, device: int = 0):
        self._lib = lib()
        h = _vp()
        rc = self._lib.frb_ctx_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise NativeError(
                f"frb_ctx_create(device={device}) failed with code {rc}: an sm_100 (B200) GPU is required; "
                "this package has no CPU or other-GPU fallback")
        self.handle = h
        self.device = int(device)
        # Wrappers that share this ctx (embedders, galleries) hold the lock across "is my upload still resident?" +
        # the native call that relies on it, so another thread cannot swap the weights / gallery in between.
        self.lock = threading.RLock()

    def gallery_generation(self) -> int:
        """Bumped by every frb_gallery_upload*: compare with the value read after one's own upload."""
        return int(self._lib.frb_gallery_generation(self.handle))

    def backbone_generation(self) -> int:
        return int(self._lib.frb_backbone_generation(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.frb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.frb_last_error(self.handle)
            raise NativeError(f"{what} failed: {msg.decode() if msg else rc}")

    def launch_count(self) -> int:
        return int(self._lib.frb_launch_count(self.handle))

    def __getattr__(self, name):
        # ctx.frb_xxx(args...) -> lib.frb_xxx(handle, args...) with error checking
        if name.startswith("frb_"):
            fn = getattr(self._lib, name)

            def call(*args):
                rc = fn(self.handle, *args)
                if fn.restype is _i:
                    self.check(rc, name)
                return rc

            return call
        raise AttributeError(name)


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
