"""Constant tables handed to the kernels (layout: include/mslesseg.h, MSL_TAB_*).

They are parameters of the path, computed once on the host with the reference's own NumPy
expressions (GC: utils/mejora_imagen.py:146; LT: :173-182; gray colormap: matplotlib
LinearSegmentedColormap 'gray' behind scripts/extraer_dataset.py:192) or extracted from OpenCV
(LUT_L / LUT_OUT: the gray <-> Lab round trip of utils/mejora_imagen.py:98-115 followed by the
BGR2GRAY of utils/utils.py:421-427; tests/test_host_logic.py
re-derives both from cv2 when it is importable).
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

TAB_LUT_L, TAB_LUT_OUT, TAB_GC, TAB_CM, TAB_LT = 0, 256, 512, 768, 1024
TABLES_BYTES = 1024 + 65536

# gray v -> L channel of cv2 GRAY2BGR + BGR2LAB (a = b = 128 for every gray)
LUT_L = np.array([
    0, 1, 1, 2, 2, 3, 5, 5, 6, 7, 7, 8, 9, 9, 10, 11, 12, 12, 14, 15, 16, 17, 18, 19, 21, 23, 24, 25, 27, 27,
    28, 30, 31, 33, 34, 35, 36, 38, 39, 40, 41, 42, 43, 45, 46, 47, 48, 50, 51, 52, 53, 54, 55, 57, 58, 59,
    60, 61, 62, 63, 65, 66, 67, 68, 69, 70, 71, 73, 74, 75, 76, 77, 78, 79, 80, 82, 82, 83, 85, 86, 87, 88,
    89, 90, 91, 92, 93, 94, 95, 97, 98, 99, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112,
    113, 114, 115, 116, 117, 119, 119, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 132, 133, 134,
    135, 136, 137, 138, 139, 140, 141, 142, 143, 144, 145, 146, 147, 148, 149, 150, 151, 152, 153, 154, 155,
    156, 156, 157, 158, 159, 160, 161, 162, 163, 164, 165, 166, 167, 168, 169, 170, 171, 172, 173, 174, 175,
    176, 177, 178, 179, 180, 180, 181, 182, 183, 184, 185, 186, 187, 188, 189, 190, 191, 192, 193, 194, 195,
    196, 196, 197, 198, 199, 200, 201, 202, 203, 204, 205, 206, 207, 208, 208, 209, 210, 211, 212, 213, 214,
    215, 216, 217, 218, 219, 219, 220, 221, 222, 223, 224, 225, 226, 227, 228, 228, 229, 230, 231, 232, 233,
    234, 235, 236, 237, 237, 238, 239, 240, 241, 242, 243, 244, 245, 245, 246, 247, 248, 249, 250, 251, 252,
    253, 253, 254, 255
], dtype=np.uint8)

# L' -> BGR2GRAY(LAB2BGR(L', 128, 128))
LUT_OUT = np.array([
    0, 2, 3, 4, 6, 7, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 23, 24, 25, 25, 26, 27, 28, 29,
    29, 30, 31, 32, 33, 34, 34, 35, 36, 37, 38, 38, 39, 40, 41, 42, 43, 43, 44, 45, 46, 47, 48, 48, 49, 50,
    51, 52, 52, 53, 54, 55, 56, 57, 58, 59, 59, 60, 61, 62, 63, 64, 65, 66, 67, 67, 68, 69, 70, 71, 72, 73,
    74, 75, 76, 77, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 90, 91, 92, 93, 94, 95, 96, 97,
    98, 99, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 112, 114, 114, 115, 116, 117,
    118, 119, 120, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 132, 133, 134, 135, 136, 137, 138,
    139, 140, 141, 143, 144, 145, 146, 147, 148, 149, 150, 151, 152, 153, 154, 155, 156, 157, 158, 159, 160,
    161, 162, 163, 164, 165, 166, 167, 168, 169, 171, 172, 173, 174, 175, 176, 177, 178, 179, 180, 181, 182,
    183, 184, 185, 186, 188, 189, 190, 191, 192, 193, 194, 195, 196, 197, 198, 199, 200, 202, 203, 204, 205,
    206, 207, 208, 209, 210, 211, 213, 214, 215, 216, 217, 218, 219, 220, 221, 222, 224, 225, 226, 227, 228,
    229, 230, 231, 232, 234, 235, 236, 237, 238, 239, 240, 241, 243, 244, 245, 246, 247, 248, 249, 250, 252,
    253, 254, 255
], dtype=np.uint8)


def gc_table(gamma: float = 2.0) -> np.ndarray:
    """reference utils/mejora_imagen.py:146 (float64, C cast)."""
    return np.array((np.linspace(0, 1, 256) ** gamma) * 255, dtype=np.uint8)


def lt_table_for_max(maxval: int) -> np.ndarray:
    """reference utils/mejora_imagen.py:173-182 on the 256 uint8 inputs of a slice whose maximum is
    `maxval`: uint16 promotes np.log to float32; NaN / inf entries (only reachable for maxval == 0,
    where the sole input value is 0 and inf * 0 = NaN casts to 0 on x86) are stored as 0."""
    img = np.arange(256, dtype=np.uint16)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = 255 / np.log(1 + np.uint16(maxval))
        out = np.clip(c * np.log(1 + img), 0, 255)
    res = np.zeros(256, dtype=np.uint8)
    ok = np.isfinite(out)
    res[ok] = out[ok].astype(np.uint8)
    return res


def gray_cmap_bytes() -> np.ndarray:
    """matplotlib 'gray' colormap as bytes (restated: lut = linspace(0, 1, 256); bytes = (lut*255) cast)."""
    return (np.linspace(0, 1, 256) * 255).astype(np.uint8)


@lru_cache(maxsize=1)
def host_tables() -> np.ndarray:
    t = np.zeros(TABLES_BYTES, dtype=np.uint8)
    t[TAB_LUT_L:TAB_LUT_L + 256] = LUT_L
    t[TAB_LUT_OUT:TAB_LUT_OUT + 256] = LUT_OUT
    t[TAB_GC:TAB_GC + 256] = gc_table()
    t[TAB_CM:TAB_CM + 256] = gray_cmap_bytes()
    for m in range(256):
        t[TAB_LT + 256 * m:TAB_LT + 256 * (m + 1)] = lt_table_for_max(m)
    t.setflags(write=False)
    return t
