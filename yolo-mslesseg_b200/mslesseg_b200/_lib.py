"""ctypes binding of libmslesseg.so (C ABI declared in include/mslesseg.h).

There is deliberately NO fallback: if the shared library is missing or a call fails the
caller gets an exception - the product path never routes through NumPy/PyTorch arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("MSLESSEG_LIB", _HERE / "libmslesseg.so"))

# enums of include/mslesseg.h
AXIAL, CORONAL, SAGITAL = 0, 1, 2
MEJORA_NONE, MEJORA_HE, MEJORA_CLAHE, MEJORA_GC, MEJORA_LT = 0, 1, 2, 3, 4
F32, U8 = 0, 1
OUT_G, OUT_P, OUT_PNG_GRAY, OUT_PNG_RGBA = 0, 1, 2, 3
WS_ENHANCE_VOLUMES, WS_RECON = 1, 2
OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
Z_RAW, Z_ZLIB, Z_GZIP, Z_PNG = 0, 1, 2, 3

PLANO_ID = {"axial": AXIAL, "coronal": CORONAL, "sagital": SAGITAL}
MEJORA_ID = {None: MEJORA_NONE, "HE": MEJORA_HE, "CLAHE": MEJORA_CLAHE, "GC": MEJORA_GC, "LT": MEJORA_LT}


class MslError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmslesseg error {code}: {msg}")
        self.code = code
        self.msg = msg


_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t

_SIGNATURES = {
    "msl_version": (C.c_int, []),
    "msl_last_error": (C.c_char_p, []),
    "msl_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i, _i]),
    "msl_lesion_slices": (C.c_int, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "msl_slice_ranges": (C.c_int, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "msl_enhance_slices": (C.c_int, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _sz, _i, _vp, _vp]),
    "msl_enhance_images": (C.c_int, [_vp, _i, _i, _i, _i, _sz, _i, _vp, _sz, _i, _vp, _vp]),
    "msl_enhance_volumes": (C.c_int, [_vp, _i, _i, _i, _i, C.POINTER(C.c_void_p), _vp, _vp, _sz, _vp]),
    "msl_png_bytes": (_sz, [_i, _i, _i]),
    "msl_png_pack": (C.c_int, [_vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "msl_deflate_bound": (_sz, [_i, _i, _sz]),
    "msl_deflate_workspace_bytes": (_sz, [_i, _i, _sz]),
    "msl_deflate_chunks": (C.c_int, [_vp, _sz, _sz, _i, _i, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "msl_deflate_files": (C.c_int, [_vp, _i, _sz, _sz, _vp, _sz, _sz, _i, _sz, _i, _i, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "msl_png_encode": (C.c_int, [_vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _sz, _vp]),
    "msl_inflate": (C.c_int, [_vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "msl_png_unfilter": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "msl_nifti_convert": (C.c_int, [_vp, _i, C.c_uint64, C.c_double, C.c_double, _i, _vp, _vp, _vp, _vp, _vp]),
    "msl_mask_contours": (C.c_int, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "msl_nonzero_flags": (C.c_int, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "msl_copy_box_d2h": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "msl_copy_boxes_d2h": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "msl_bgr_to_gray": (C.c_int, [_vp, _sz, _vp, _vp]),
    "msl_combine_predictions": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "msl_recon": (C.c_int, [_vp, _sz, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "msl_consensus_eval": (C.c_int, [_vp, _vp, _vp, _vp, _i, _sz, _i, _vp, _vp, _vp]),
    "msl_confusion_counts": (C.c_int, [_vp, _vp, _i, _sz, _vp, _vp]),
    "msl_slice_counts": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "msl_kernel_kinds": (C.c_int, []),
    "msl_kernel_name": (C.c_char_p, [_i]),
    "msl_kernel_launches": (C.c_ulonglong, [C.POINTER(C.c_ulonglong)]),
    "msl_profile_enable": (C.c_int, [_i]),
    "msl_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "msl_stage_slices": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "msl_enhance_stack_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "msl_enhance_stack": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msl_selftest_norm_division": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "msl_profile_timeline": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libmslesseg.so once; raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} not found - build it with `make -C yolo-mslesseg_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc: int) -> None:
    if rc != OK:
        msg = load().msl_last_error()
        raise MslError(rc, msg.decode("utf-8", "replace") if msg else "")


def kernel_launches() -> dict:
    """{kernel kind name: launches since the library was loaded}."""
    lib = load()
    n = lib.msl_kernel_kinds()
    arr = (C.c_ulonglong * n)()
    lib.msl_kernel_launches(arr)
    return {lib.msl_kernel_name(k).decode(): int(arr[k]) for k in range(n)}


def profile_enable(on: bool = True) -> None:
    check(load().msl_profile_enable(1 if on else 0))


def profile_timeline(cap: int = 1 << 16) -> list:
    """[(kernel kind name, stream handle, start ms, end ms)] of the launches recorded since profile_enable(), in launch
    order; call it before profile_collect()."""
    lib = load()
    kind = (C.c_int * cap)(); stream = (C.c_ulonglong * cap)(); t0 = (C.c_double * cap)(); t1 = (C.c_double * cap)()
    n = lib.msl_profile_timeline(cap, kind, stream, t0, t1)
    if n < 0:
        raise RuntimeError(lib.msl_last_error().decode())
    return [(lib.msl_kernel_name(kind[i]).decode(), int(stream[i]), float(t0[i]), float(t1[i])) for i in range(n)]


def profile_collect() -> dict:
    """{kernel kind name: (milliseconds, launches)} for the launches since profile_enable()."""
    lib = load()
    n = lib.msl_kernel_kinds()
    ms = (C.c_double * n)()
    cnt = (C.c_ulonglong * n)()
    check(lib.msl_profile_collect(ms, cnt))
    return {lib.msl_kernel_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n) if cnt[k]}
