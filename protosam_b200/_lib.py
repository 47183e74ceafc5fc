"""ctypes binding of libpsam_b200.so (include/psam_b200.h).

The library is the product: there is no Python/PyTorch fallback.  ``load()`` raises if the
shared object has not been built (``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C protosam_b200/csrc``) and every wrapper raises ``RuntimeError`` with the library's own
message when a call is rejected.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpsam_b200.so")

MODE_MASK, MODE_GRIDCONV, MODE_GRIDCONV_PLUS, MODE_AUTO_FG = 0, 1, 2, 3
MODE_IDS = {"mask": MODE_MASK, "gridconv": MODE_GRIDCONV, "gridconv+": MODE_GRIDCONV_PLUS, "auto_fg": MODE_AUTO_FG}
MODE_NAMES = {v: k for k, v in MODE_IDS.items()}

SET_EMPTY = 1
ABI_VERSION = 3
PROB_SOFTMAX, PROB_SOFTMAX_TWICE = 0, 1
IMG_EMPTY, IMG_RUN_OVERFLOW, IMG_CC_TRUNCATED, IMG_CCA_AMBIGUOUS = 1, 2, 4, 8
REC_SELECTED = 1
PACKED_OVERFLOW = 1

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t
c_f = ctypes.c_float

# name -> (restype, argtypes); mirrors include/psam_b200.h one to one
SIGNATURES = {
    "psam_abi_version": (c_i, []),
    "psam_last_error": (ctypes.c_char_p, []),
    "psam_launch_count": (ctypes.c_uint64, []),
    "psam_trace_install": (c_i, [c_p, c_sz]),
    "psam_profile_enable": (None, [c_i]),
    "psam_profile_collect": (c_i, [c_p, c_sz, c_p, c_p, c_i]),
    "psam_alp_prototypes_workspace": (c_sz, [c_i] * 7),
    "psam_alp_prototypes": (c_i, [c_p, c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f,
                                  c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "psam_alp_prototypes_shots": (c_i, [c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f,
                                        c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "psam_combine_shots": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "psam_tokens_to_features": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "psam_mask_nearest": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "psam_alp_proto_grid": (c_i, [c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_p]),
    "psam_alp_match_workspace": (c_sz, [c_i] * 6),
    "psam_match_reserve_sms": (c_i, [c_i]),
    "psam_alp_match": (c_i, [c_p, c_i64, c_i64, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p,
                             c_p, c_sz, c_i, c_p]),
    "psam_upsample_workspace": (c_sz, [c_i, c_i]),
    "psam_upsample_softmax": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_sz, c_p]),
    "psam_prompts_workspace": (c_sz, [c_i] * 4),
    "psam_components": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "psam_records_to_sam": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "psam_coarse_to_prompts_workspace": (c_sz, [c_i] * 4),
    "psam_coarse_to_prompts": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_sz, c_p]),
    "psam_packed_bytes": (c_sz, [c_i, c_i]),
    "psam_compact_records": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "psam_peer_region_bytes": (c_sz, [c_sz]),
    "psam_peer_push_table": (c_i, [c_p, c_i, c_i, c_i, c_sz, c_sz, c_p, c_i, c_i, c_p, c_p]),
    "psam_peer_recv_table": (c_i, [c_p, c_i, c_i, c_i, c_sz, c_sz, c_p, c_i, c_i, c_i, c_p, c_p]),
    "psam_peer_put": (c_i, [c_p, c_sz, c_sz, c_p, c_i, c_i, c_i, c_p, c_p]),
    "psam_peer_collect": (c_i, [c_p, c_sz, c_p, c_i, c_i, c_p, c_p, c_p]),
    "psam_topk_points_workspace": (c_sz, [c_i] * 3),
    "psam_topk_points": (c_i, [c_p, c_p, c_i64, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_sz, c_p]),
    "psam_neg_points": (c_i, [c_p, c_p, c_i64, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_p]),
    "psam_mask_prompts": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "psam_confidence": (c_i, [c_p, c_i, c_i64, c_p, c_p]),
}

_lib = None


def load():
    """Load the CUDA library or fail loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"protosam_b200: CUDA library not built ({LIB_PATH} missing). Build it with "
            "`make -C protosam_b200/csrc` (nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.psam_abi_version() != ABI_VERSION:
        raise RuntimeError("protosam_b200: ABI version mismatch between _lib.py and libpsam_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().psam_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def match_reserve_sms(n: int = -1) -> int:
    """SMs the persistent match kernels leave free (multi-GPU runs: room for the NCCL kernels of other volumes in
    flight); n < 0 only queries.  Returns the previous value."""
    return int(load().psam_match_reserve_sms(int(n)))


def launch_count() -> int:
    return int(load().psam_launch_count())


def profile_enable(on: bool):
    load().psam_profile_enable(1 if on else 0)


def profile_collect(max_kernels: int = 64) -> dict:
    """-> {kernel name: (total ms, launches)} since the last collect (synchronises on the recorded events)."""
    names = ctypes.create_string_buffer(64 * max_kernels)
    ms = (ctypes.c_float * max_kernels)()
    cnt = (ctypes.c_int32 * max_kernels)()
    n = load().psam_profile_collect(names, len(names), ms, cnt, max_kernels)
    keys = names.value.decode().split("\n")[:n]
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(keys)}
