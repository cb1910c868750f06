"""ctypes binding of libcfa_b200.so (C ABI declared in include/cfa_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing the import fails, and
every compute entry point raises unless it runs on a CUDA (sm_100a) device.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libcfa_b200.so")

DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class CfaError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m clip_finegrained_alignment_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU/PyTorch fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

SIGNATURES = {
    "cfa_abi_version": (C.c_int, []),
    "cfa_error_string": (C.c_char_p, [_i]),
    "cfa_adamspd_chunk_elems": (C.c_int, []),
    "cfa_adamspd_step": (C.c_int, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "cfa_adamspd_step_amp": (C.c_int, [_vp, _i, _vp, _i, _vp, _vp, _vp, _f, _vp, _i, _i, _vp]),
    "cfa_global_infonce_workspace_bytes": (_sz, [_i, _i, _i]),
    "cfa_global_infonce_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _i, _f, _f,
                                         _vp, _vp, _sz, _i, _i, _vp]),
    "cfa_global_infonce_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _sz, _i, _i, _vp]),
    "cfa_global_infonce_path": (C.c_int, [_i, _i, _i, _i]),
    "cfa_sparc_fwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _vp, _vp, _sz, _i, _vp]),
    "cfa_sparc_bwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "cfa_sparc_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "cfa_sparc_loss_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "cfa_sparc_loss_fwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _f, _vp, _sz, _i, _vp]),
    "cfa_sparc_loss_bwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _f, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _i, _vp]),
    "cfa_peer_exchange_bytes": (_sz, [_i, _i]),
    "cfa_peer_alloc": (C.c_int, [_sz, _vp, _vp]),
    "cfa_peer_open": (C.c_int, [_vp, _vp]),
    "cfa_peer_close": (C.c_int, [_vp]),
    "cfa_peer_free": (C.c_int, [_vp]),
    "cfa_peer_sync": (C.c_int, [_vp, _i, _i, C.c_uint32, _vp, _sz, _sz, _sz, _i, _vp, _vp]),
    "cfa_peer_status": (C.c_int, [_vp, _vp, _vp]),
    "cfa_sparc_loss_gathered_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "cfa_sparc_loss_gathered_fwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _f, _vp, _sz, _i, _i, _i, _vp,
                                              C.c_uint32, _vp]),
    "cfa_sparc_loss_gathered_bwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _f, _vp, _sz, _vp, _vp, _vp, _vp,
                                              _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "cfa_sparc_loss_gathered_bwd_ex": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _f, _vp, _sz, _vp, _vp, _vp, _vp,
                                                 _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "cfa_count_contrastive_fwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp]),
    "cfa_count_contrastive_bwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "cfa_logits_ce_fwd": (C.c_int, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "cfa_logits_ce_bwd": (C.c_int, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cfa_global_infonce_gathered_workspace_bytes": (_sz, [_i, _i, _i]),
    "cfa_global_infonce_gathered_fwd": (C.c_int, [_vp, _i, _i, _f, _f, _vp, _sz, _i, _i, _vp, C.c_uint32, _vp, _vp]),
    "cfa_global_infonce_gathered_bwd": (C.c_int, [_vp, _i, _i, _f, _f, _vp, _sz, _vp, _vp, _i, _i, _vp]),
    "cfa_masked_pairwise_fwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "cfa_masked_pairwise_bwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cfa_sparc_path": (C.c_int, [_i, _i, _i, _i, _i]),
    "cfa_sparc_bwd_path": (C.c_int, [_i, _i, _i, _i, _i]),
    "cfa_sparc_max_patches": (C.c_int, [_i, _i]),
    "cfa_sparc_finalize": (C.c_int, [_vp, _i, _vp, _vp, _i, _i, _f, _f, _vp, _i, _vp]),
    "cfa_sparc_coef": (C.c_int, [_vp, _f, _f, _i, _vp, _vp, _vp]),
    "cfa_sparc_coef_ptrs": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp]),
    "cfa_debug_set_profile_buffer": (C.c_int, [_vp]),
    "cfa_debug_set_profile_buffer_fwd": (C.c_int, [_vp]),
    "cfa_debug_set_marker_buffer": (C.c_int, [_vp]),
    "cfa_tc_selftest": (C.c_int, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "cfa_tc_selftest_timed": (C.c_int, [_i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


# kernels launched per C-ABI call (cudaMemsetAsync not counted); bench.py reports the running total
LAUNCHES = {"cfa_adamspd_step": 2, "cfa_adamspd_step_amp": 4, "cfa_global_infonce_fwd": 1, "cfa_global_infonce_bwd": 2,
            "cfa_sparc_fwd": 1, "cfa_sparc_bwd": 1, "cfa_sparc_finalize": 1, "cfa_sparc_coef": 1, "cfa_sparc_coef_ptrs": 1,
            "cfa_masked_pairwise_fwd": 2, "cfa_masked_pairwise_bwd": 1,
            "cfa_sparc_loss_fwd": 4, "cfa_sparc_loss_bwd": 3,     # tensor-core chain: fwd3 + split + logits + combine | logits-bwd + norm-bwd + bwd3 (coefficients evaluated in-kernel)
            "cfa_count_contrastive_fwd": 2, "cfa_count_contrastive_bwd": 1, "cfa_logits_ce_fwd": 2, "cfa_logits_ce_bwd": 1,
            "cfa_sparc_loss_gathered_fwd": 8, "cfa_sparc_loss_gathered_bwd": 4, "cfa_sparc_loss_gathered_bwd_ex": 4, "cfa_peer_sync": 1,
            "cfa_global_infonce_gathered_fwd": 7, "cfa_global_infonce_gathered_bwd": 2,
            "cfa_tc_selftest": 1, "cfa_tc_selftest_timed": 1}
launch_count = 0
kernel_events = None       # {abi name: [(start_event, end_event), ...]} while bench.py profiles; else None


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point on the current stream, raising CfaError on a non-zero status."""
    global launch_count
    ev = kernel_events.get(name) if kernel_events is not None else None
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        ev.append((e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    launch_count += LAUNCHES.get(name, 0)
    check(rc, name)


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib.cfa_error_string(code)
        raise CfaError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    # torch.cuda.current_stream() costs ~20 us of Python per call; the raw query is ~1 us
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise CfaError("clip_finegrained_alignment_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise CfaError("all tensors of one call must live on the same device")
    return dev
