"""ctypes bindings of the C ABI declared in include/oarfish_em.h.

The CUDA library is built in-tree (``make product`` or ``__graft_entry__.build()``)
as ``oarfish_b200/lib/liboarfish_em.so``.  There is no CPU fallback: if the
library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
# OAR_EM_LIB points at another build of the same library (kernel-tuning variants built by tools/dev/build_variants.sh)
EM_LIB_PATH = os.environ.get("OAR_EM_LIB") or os.path.join(LIB_DIR, "liboarfish_em.so")
SYNTH_LIB_PATH = os.path.join(LIB_DIR, "liboarsynth.so")

OAR_OK = 0
OAR_ERR_INVALID = -1
OAR_ERR_CUDA = -2
OAR_ERR_OOM = -3
OAR_ERR_UNSUPPORTED = -4

KERNEL_AUTO = 0
KERNEL_ROWGROUP = 1
KERNEL_TILED = 2

# every symbol include/oarfish_em.h declares: name -> (restype, argtypes)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

ABI = {
    "oar_version": (C.c_int, []),
    "oar_device_count": (C.c_int, []),
    "oar_last_error": (C.c_char_p, []),
    "oar_store_create": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(_vp)]),
    "oar_store_destroy": (None, [_vp]),
    "oar_store_info": (C.c_int, [_vp, _u64p, _u64p, _u32p, C.POINTER(C.c_int)]),
    "oar_store_set_kernel": (C.c_int, [_vp, C.c_int]),
    "oar_em": (C.c_int, [_vp, _vp, C.c_uint32, C.c_double, C.c_uint32, _vp, _u32p, _f64p]),
    "oar_bootstrap": (C.c_int, [_vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, _vp, _vp]),
    "oar_bootstrap_weights": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, _vp, _vp]),
    "oar_bootstrap_sample_weights": (C.c_int, [_vp, C.c_uint64, C.c_uint32, _vp]),
    "oar_em_batched": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, _vp, _vp, _vp, C.c_uint64, _u64p, _vp]),
    "oar_store_coverage_model": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_double, _vp]),
    "oar_store_coverage_model_binomial": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp]),
    "oar_posteriors": (C.c_int, [_vp, _vp, C.c_double, _vp, _vp]),
    "oar_aux_counts": (C.c_int, [_vp, _vp, _vp]),
    "oar_store_layout_info": (C.c_int, [_vp, _u64p]),
    "oar_store_layout_lpos": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp]),
    "oar_store_timings": (C.c_int, [_vp, _f64p]),
    "oar_store_counters": (C.c_int, [_vp, _u64p]),
    "oar_sweep": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "oar_sweep_timed": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_float)]),
    "oar_store_stream": (_vp, [_vp]),
    "oar_store_create_filtered": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, C.c_uint32, _vp, C.c_int,
                                             C.POINTER(_vp), _vp, _vp, _vp]),
    "oar_store_export": (C.c_int, [_vp, _vp, _vp, _vp]),
    "oar_store_set_progress": (C.c_int, [_vp, _vp, _vp]),
    "oar_multi_create": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, _vp, C.c_int, C.POINTER(_vp)]),
    "oar_multi_destroy": (None, [_vp]),
    "oar_multi_store": (_vp, [_vp, C.c_int]),
    "oar_multi_info": (C.c_int, [_vp, C.POINTER(C.c_int), _f64p, _vp]),
    "oar_multi_bootstrap": (C.c_int, [_vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_double, _vp, _vp]),
    "oar_em_batched_multi": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, _vp, C.c_uint32, _vp, C.c_int,
                                        C.c_uint32, C.c_double, C.c_uint32, _vp, _vp, _vp, C.c_uint64, _u64p, _vp, _vp]),
}

class FilterOpts(C.Structure):
    """oar_filter_opts (include/oarfish_em.h) == the fields of AlignmentFilters that filter() reads."""
    _fields_ = [("which_strand", C.c_int32), ("min_aligned_len", C.c_uint32), ("three_prime_clip", C.c_int64),
                ("five_prime_clip", C.c_uint32), ("min_aligned_fraction", C.c_float), ("score_threshold", C.c_float),
                ("score_prob_denom", C.c_float)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_uint32, C.c_double, C.c_void_p)

_em_lib = None
_synth_lib = None


class OarfishError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oarfish_em error {code}: {msg}")
        self.code = code


def load_em_lib() -> C.CDLL:
    """Load liboarfish_em.so and bind every ABI symbol.  Raises if absent."""
    global _em_lib
    if _em_lib is not None:
        return _em_lib
    if not os.path.exists(EM_LIB_PATH):
        raise ImportError(
            f"{EM_LIB_PATH} not found: build it with `make product` (or __graft_entry__.build()). "
            "oarfish_b200 has no CPU fallback.")
    lib = C.CDLL(EM_LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in ABI.items():
        if os.environ.get("OAR_EM_LIB") and not hasattr(lib, name):
            continue  # an older build selected for A/B timing (tools/dev): entry points added since are simply absent
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _em_lib = lib
    return lib


def load_synth_lib() -> C.CDLL:
    global _synth_lib
    if _synth_lib is not None:
        return _synth_lib
    if not os.path.exists(SYNTH_LIB_PATH):
        raise ImportError(f"{SYNTH_LIB_PATH} not found: build it with `make product`.")
    lib = C.CDLL(SYNTH_LIB_PATH)
    lib.oar_synth_plan.restype = C.c_int
    lib.oar_synth_plan.argtypes = [C.c_uint64, C.c_uint32, C.c_double, C.c_uint64, _vp]
    lib.oar_synth_fill.restype = C.c_int
    lib.oar_synth_fill.argtypes = [C.c_uint64, C.c_uint32, C.c_double, C.c_uint64, _vp, _vp, _vp, _vp, _vp]
    _synth_lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OAR_OK:
        msg = load_em_lib().oar_last_error()
        raise OarfishError(rc, msg.decode("utf-8", "replace") if msg else "")
