"""CPU ORACLE for the YOLO-MSLesSeg voxel path.  TEST INFRASTRUCTURE - NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module; the product (yolo-mslesseg_b200/) never does and has no CPU
fallback.

What this is: a NumPy restatement of the arithmetic the reference performs on the hot
path.  Every function cites the reference
file:line it follows (paths relative to the reference checkout).  Where the reference
delegates to OpenCV (third-party, opencv-python pinned 4.11.0.86 in requirements.txt:25,
source not vendored) the published algorithm is restated in NumPy: `equalizeHist`
(imgproc/histogram.cpp) and `CLAHE::apply` (imgproc/clahe.cpp); the two colour-space
round trips collapse to the 256-entry tables LUT_L / LUT_OUT.

Parity status
-------------
The reference has NO tests or golden vectors of its own, so parity is
pinned the other way the task allows: `oracle/make_golden.py` imports the real reference
in the build container, runs it on the two demo volumes and on seeded synthetic volumes
and freezes the outputs under tests/golden/.  tests/test_oracle_golden.py checks this
file against those fixtures (and against cv2 itself when cv2 is importable).
Row E8 (`matplotlib.pyplot.imsave`) is restated from matplotlib 3.10 sources from memory;
matplotlib is not installed here -> **E8 parity is unpinned**.

Array conventions: volumes are NumPy arrays shaped (X, Y, Z) exactly as the reference
sees them after `nib.load().get_fdata()`; `vol_xyz = dev[z, y, x].transpose(2, 1, 0)` maps
a C-contiguous [Z][Y][X] device buffer to it without a copy.
"""
from __future__ import annotations

import numpy as np

PLANOS = ("axial", "coronal", "sagital")
MEJORAS = ("HE", "CLAHE", "GC", "LT")

# --------------------------------------------------------------------------------------
# Constant tables
# --------------------------------------------------------------------------------------

# gray v -> L of cv2.cvtColor(GRAY2BGR -> BGR2LAB); a = b = 128 for every gray
# (reference call sites: utils/mejora_imagen.py:98,101).  Extracted from cv2 4.13.0.
LUT_L = np.array([
    0, 1, 1, 2, 2, 3, 5, 5, 6, 7, 7, 8, 9, 9, 10, 11, 12, 12, 14, 15, 16, 17, 18, 19, 21, 23, 24, 25, 27, 27, 28, 30,
    31, 33, 34, 35, 36, 38, 39, 40, 41, 42, 43, 45, 46, 47, 48, 50, 51, 52, 53, 54, 55, 57, 58, 59, 60, 61, 62, 63, 65, 66, 67, 68,
    69, 70, 71, 73, 74, 75, 76, 77, 78, 79, 80, 82, 82, 83, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95, 97, 98, 99, 100, 101, 102, 103,
    104, 105, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 116, 117, 119, 119, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 132, 133, 134, 135, 136,
    137, 138, 139, 140, 141, 142, 143, 144, 145, 146, 147, 148, 149, 150, 151, 152, 153, 154, 155, 156, 156, 157, 158, 159, 160, 161, 162, 163, 164, 165, 166, 167,
    168, 169, 170, 171, 172, 173, 174, 175, 176, 177, 178, 179, 180, 180, 181, 182, 183, 184, 185, 186, 187, 188, 189, 190, 191, 192, 193, 194,
    195, 196, 196, 197, 198, 199, 200, 201, 202, 203, 204, 205, 206, 207, 208, 208, 209, 210, 211, 212, 213, 214, 215, 216, 217, 218, 219, 219, 220, 221, 222, 223,
    224, 225, 226, 227, 228, 228, 229, 230, 231, 232, 233, 234, 235, 236, 237, 237, 238, 239, 240, 241, 242, 243, 244, 245, 245, 246, 247, 248, 249, 250, 251, 252,
    253, 253, 254, 255], dtype=np.uint8)

# L' -> cv2.cvtColor(LAB2BGR of (L',128,128)) -> BGR2GRAY, i.e. what
# utils/mejora_imagen.py:112-115 followed by utils/utils.py:421-427 does to the CLAHE output.
LUT_OUT = np.array([
    0, 2, 3, 4, 6, 7, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 23, 24, 25, 25, 26, 27, 28, 29, 29, 30, 31, 32,
    33, 34, 34, 35, 36, 37, 38, 38, 39, 40, 41, 42, 43, 43, 44, 45, 46, 47, 48, 48, 49, 50, 51, 52, 52, 53, 54, 55, 56, 57, 58, 59,
    59, 60, 61, 62, 63, 64, 65, 66, 67, 67, 68, 69, 70, 71, 72, 73, 74, 75, 76, 77, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88,
    89, 90, 90, 91, 92, 93, 94, 95, 96, 97, 98, 99, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 112, 114, 114, 115, 116, 117, 118,
    119, 120, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 132, 133, 134, 135, 136, 137, 138, 139, 140, 141, 143, 144, 145, 146, 147, 148, 149, 150, 151,
    152, 153, 154, 155, 156, 157, 158, 159, 160, 161, 162, 163, 164, 165, 166, 167, 168, 169, 171, 172, 173, 174, 175, 176, 177, 178, 179, 180, 181, 182, 183, 184,
    185, 186, 188, 189, 190, 191, 192, 193, 194, 195, 196, 197, 198, 199, 200, 202, 203, 204, 205, 206, 207, 208, 209, 210, 211, 213, 214, 215, 216, 217, 218, 219,
    220, 221, 222, 224, 225, 226, 227, 228, 229, 230, 231, 232, 234, 235, 236, 237, 238, 239, 240, 241, 243, 244, 245, 246, 247, 248, 249, 250, 252, 253, 254, 255],
    dtype=np.uint8)


def gc_table(gamma: float = 2.0) -> np.ndarray:
    """utils/mejora_imagen.py:146 - float64 linspace ** gamma * 255, C cast (truncation)."""
    return np.array((np.linspace(0, 1, 256) ** gamma) * 255, dtype=np.uint8)


def lt_table(maxval: int = 255) -> np.ndarray:
    """utils/mejora_imagen.py:173-182 evaluated on the 256 possible uint8 inputs for a slice
    whose maximum is `maxval`.  uint16 input makes NumPy pick float32 for `np.log` (NEP 50),
    so `c` and the product are float32.  maxval == 0 gives c = inf, inf*0 = NaN and the x86
    NaN -> uint8 cast yields 0."""
    img = np.arange(256, dtype=np.uint16)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = 255 / np.log(1 + np.uint16(maxval))
        out = np.clip(c * np.log(1 + img), 0, 255)
    res = np.zeros(256, dtype=np.uint8)
    ok = np.isfinite(out)
    res[ok] = out[ok].astype(np.uint8)
    return res


def gray_cmap_bytes() -> np.ndarray:
    """matplotlib 'gray' colormap as bytes [RESTATED, unpinned]: LinearSegmentedColormap LUT =
    linspace(0,1,256); Colormap.__call__(bytes=True) uses (lut*255).astype(uint8)."""
    return (np.linspace(0, 1, 256) * 255).astype(np.uint8)


# --------------------------------------------------------------------------------------
# E1  float -> uint8 per-slice normalisation
# --------------------------------------------------------------------------------------

def normalizar_a_uint8(imagen: np.ndarray) -> np.ndarray:
    """utils/utils.py:396-406.  float32 throughout: sub(min), ptp, div THEN mul(255), trunc."""
    if imagen.dtype == np.uint8:
        return imagen
    f = imagen.astype(np.float32)
    f = f - np.min(f)
    p = np.ptp(f)
    if p > 0:
        f = np.float32(255) * (f / p)
    return f.astype(np.uint8)


# --------------------------------------------------------------------------------------
# E3  HE == cv2.equalizeHist on the normalised slice
# --------------------------------------------------------------------------------------

def equalize_hist(u: np.ndarray) -> np.ndarray:
    """OpenCV imgproc/histogram.cpp `equalizeHist` (call site utils/mejora_imagen.py:62).
    SURVEY Appendix A.3."""
    hist = np.bincount(u.ravel(), minlength=256).astype(np.int64)
    total = int(u.size)
    i0 = int(np.nonzero(hist)[0][0])
    if hist[i0] == total:
        return np.full_like(u, i0)
    scale = np.float32(255.0) / np.float32(total - hist[i0])
    lut = np.zeros(256, dtype=np.uint8)
    csum = np.cumsum(hist[i0 + 1:]).astype(np.float32)
    vals = np.rint(csum * scale)                       # cvRound: round-half-even, float32 product
    lut[i0 + 1:] = np.clip(vals, 0, 255).astype(np.uint8)
    return lut[u]


# --------------------------------------------------------------------------------------
# E4  CLAHE (OpenCV semantics), clipLimit 2.0, 8x8 tiles
# --------------------------------------------------------------------------------------

def clahe_geometry(rows: int, cols: int, tiles_x: int = 8, tiles_y: int = 8, clip_limit: float = 2.0):
    """imgproc/clahe.cpp CLAHE_Impl::apply: padded size, tile size, integer clip, lutScale."""
    if cols % tiles_x == 0 and rows % tiles_y == 0:
        prow, pcol = rows, cols
    else:
        prow = rows + (tiles_y - rows % tiles_y)
        pcol = cols + (tiles_x - cols % tiles_x)
    th, tw = prow // tiles_y, pcol // tiles_x
    area = th * tw
    clip = max(int(clip_limit * area / 256), 1)
    lut_scale = np.float32(255.0) / np.float32(area)
    return prow, pcol, th, tw, clip, lut_scale


def clahe_tile_luts(L: np.ndarray, tiles_x: int = 8, tiles_y: int = 8, clip_limit: float = 2.0) -> np.ndarray:
    """Per-tile LUTs, shape (tiles_y, tiles_x, 256) uint8 (CLAHE_CalcLut_Body)."""
    rows, cols = L.shape
    prow, pcol, th, tw, clip, lut_scale = clahe_geometry(rows, cols, tiles_x, tiles_y, clip_limit)
    ext = L if (prow, pcol) == (rows, cols) else np.pad(L, ((0, prow - rows), (0, pcol - cols)), mode="reflect")
    t = ext.reshape(tiles_y, th, tiles_x, tw).transpose(0, 2, 1, 3).reshape(tiles_y * tiles_x, th * tw)
    ntile = tiles_y * tiles_x
    flat = (np.arange(ntile, dtype=np.int64)[:, None] * 256 + t.astype(np.int64)).ravel()
    hist = np.bincount(flat, minlength=ntile * 256).reshape(ntile, 256).astype(np.int64)
    clipped = np.maximum(hist - clip, 0).sum(axis=1)
    hist = np.minimum(hist, clip)
    rb = clipped // 256
    res = clipped - rb * 256
    hist = hist + rb[:, None]
    for k in range(ntile):
        r = int(res[k])
        if r > 0:
            step = max(256 // r, 1)
            idx = np.arange(0, 256, step)[:r]
            hist[k, idx] += 1
    csum = np.cumsum(hist, axis=1).astype(np.float32)
    lut = np.clip(np.rint(csum * lut_scale), 0, 255).astype(np.uint8)
    return lut.reshape(tiles_y, tiles_x, 256)


def clahe_apply(L: np.ndarray, tiles_x: int = 8, tiles_y: int = 8, clip_limit: float = 2.0) -> np.ndarray:
    """cv2.createCLAHE(clip_limit, (tiles_x, tiles_y)).apply(L) restated."""
    rows, cols = L.shape
    _, _, th, tw, _, _ = clahe_geometry(rows, cols, tiles_x, tiles_y, clip_limit)
    lut = clahe_tile_luts(L, tiles_x, tiles_y, clip_limit).astype(np.float32)
    f32 = np.float32
    inv_tw = f32(1.0) / f32(tw)
    inv_th = f32(1.0) / f32(th)
    xs = np.arange(cols, dtype=np.float32)
    txf = xs * inv_tw - f32(0.5)
    tx1 = np.floor(txf).astype(np.int64)
    tx2 = tx1 + 1
    xa = (txf - tx1.astype(np.float32)).astype(np.float32)
    xa1 = (f32(1.0) - xa).astype(np.float32)
    tx1 = np.maximum(tx1, 0)
    tx2 = np.minimum(tx2, tiles_x - 1)
    ys = np.arange(rows, dtype=np.float32)
    tyf = ys * inv_th - f32(0.5)
    ty1 = np.floor(tyf).astype(np.int64)
    ty2 = ty1 + 1
    ya = (tyf - ty1.astype(np.float32)).astype(np.float32)
    ya1 = (f32(1.0) - ya).astype(np.float32)
    ty1 = np.maximum(ty1, 0)
    ty2 = np.minimum(ty2, tiles_y - 1)
    v = L.astype(np.int64)
    Y1, X1 = ty1[:, None], tx1[None, :]
    Y2, X2 = ty2[:, None], tx2[None, :]
    XA, XA1 = xa[None, :], xa1[None, :]
    YA, YA1 = ya[:, None], ya1[:, None]
    top = lut[Y1, X1, v] * XA1 + lut[Y1, X2, v] * XA
    bot = lut[Y2, X1, v] * XA1 + lut[Y2, X2, v] * XA
    res = top * YA1 + bot * YA
    assert res.dtype == np.float32
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------
# E3-E7  one enhanced gray slice  G = verificar_grises(<Mejora>().aplicar(S))
# --------------------------------------------------------------------------------------

def enhance_u8(u: np.ndarray, mejora: str) -> np.ndarray:
    """Enhancement of an already-normalised uint8 slice -> gray uint8 slice.
    HE   utils/mejora_imagen.py:52-67   (== equalizeHist(u): Y=u, U=V=128)
    CLAHE utils/mejora_imagen.py:91-117 (== LUT_OUT[clahe(LUT_L[u])])
    GC   utils/mejora_imagen.py:139-151 (== gc_table()[u])
    LT   utils/mejora_imagen.py:166-184 (== lt_table(max u)[u])
    followed by verificar_grises utils/utils.py:421-427 (BGR2GRAY; identity on grays)."""
    if mejora == "HE":
        return equalize_hist(u)
    if mejora == "CLAHE":
        return LUT_OUT[clahe_apply(LUT_L[u])]
    if mejora == "GC":
        return gc_table()[u]
    if mejora == "LT":
        return lt_table(int(u.max()))[u]
    raise ValueError(f"Mejora no reconocida: {mejora}.")


def enhance_slice(S: np.ndarray, mejora: str) -> np.ndarray:
    """E1 + enhancement: raw (float) slice -> gray uint8 slice G, shape == S.shape."""
    return enhance_u8(normalizar_a_uint8(np.asarray(S)), mejora)


# --------------------------------------------------------------------------------------
# E2 / E7  slicing
# --------------------------------------------------------------------------------------

def plane_axis(plano: str) -> int:
    """utils/Paciente.py:186-193 (num_cortes mapping)."""
    return {"axial": 2, "coronal": 1, "sagital": 0}[plano]


def slice_of(vol_xyz: np.ndarray, plano: str, i: int) -> np.ndarray:
    """utils/Paciente.py:230-246 indice_plano."""
    if plano == "axial":
        return vol_xyz[:, :, i]
    if plano == "coronal":
        return vol_xyz[:, i, :]
    if plano == "sagital":
        return vol_xyz[i, :, :]
    raise ValueError(f"Plano {plano} no válido.")


def png_orient(G: np.ndarray) -> np.ndarray:
    """scripts/extraer_dataset.py:192 - `imsave(G.T, origin="lower")` flips the rows:
    P[r, c] = G[c, cols-1-r], shape (cols, rows)."""
    return np.ascontiguousarray(G.T[::-1])


# --------------------------------------------------------------------------------------
# E8  matplotlib.pyplot.imsave(path, A, cmap="gray", origin="lower")   [RESTATED - UNPINNED]
# --------------------------------------------------------------------------------------

def imsave_gray_index(A: np.ndarray) -> np.ndarray:
    """colors.Normalize.__call__ + Colormap.__call__ index computation on the (already
    transposed) array A; float32 working dtype for integer input, float64 for float64."""
    A = np.asarray(A)
    if A.dtype.kind in "ui" or A.dtype == np.bool_:
        dt = np.promote_types(A.dtype, np.float32)
    else:
        dt = A.dtype
    vmin, vmax = dt.type(A.min()), dt.type(A.max())
    if vmin == vmax:
        return np.zeros(A.shape, dtype=np.uint8)
    t = A.astype(dt)
    t = t - vmin
    t = t / (vmax - vmin)
    t = t * dt.type(256)
    t[t == 256] = 255
    return t.astype(np.int64).clip(0, 255).astype(np.uint8)


def imsave_gray(slice_2d: np.ndarray) -> np.ndarray:
    """Gray byte image (cols, rows) that ends up in R=G=B of the saved PNG for
    `plt.imsave(path, slice_2d.T, cmap="gray", origin="lower")`."""
    A = np.asarray(slice_2d).T[::-1]
    return gray_cmap_bytes()[imsave_gray_index(A)]


def imsave_rgba(slice_2d: np.ndarray) -> np.ndarray:
    g = imsave_gray(slice_2d)
    out = np.empty(g.shape + (4,), dtype=np.uint8)
    out[..., 0] = g
    out[..., 1] = g
    out[..., 2] = g
    out[..., 3] = 255
    return out


# --------------------------------------------------------------------------------------
# E0  lesion-slice selection
# --------------------------------------------------------------------------------------

def indices_cortes_con_lesion(gt_xyz: np.ndarray, plano: str) -> list:
    """utils/Paciente.py:252-259."""
    n = gt_xyz.shape[plane_axis(plano)]
    return [i for i in range(n) if np.any(slice_of(gt_xyz, plano, i) > 0)]


def ventana_central(indices_validos: list, num_cortes) -> list:
    """utils/Paciente.py:261-275 (the list arithmetic of indices_a_usar)."""
    if num_cortes is None or len(indices_validos) <= num_cortes:
        return list(indices_validos)
    centro = len(indices_validos) // 2
    mitad = num_cortes // 2
    start = max(0, centro - mitad)
    return list(indices_validos[start:start + num_cortes])


def indices_a_usar(gt_xyz: np.ndarray, plano: str, num_cortes=None) -> list:
    return ventana_central(indices_cortes_con_lesion(gt_xyz, plano), num_cortes)


def num_cortes_percentil(conteos: list, percentil: int = 50) -> int:
    """scripts/extraer_dataset.py:110-135 - int(np.percentile(counts, p))."""
    if not conteos:
        raise ValueError("No se encontraron cortes con lesión válidos para calcular el percentil.")
    return int(np.percentile(conteos, percentil))


# --------------------------------------------------------------------------------------
# E8 container  PNG file around the imsave pixels (encode side)
# --------------------------------------------------------------------------------------

def png_stored(pixels: np.ndarray) -> bytes:
    """A complete 8-bit PNG file (colour type 6 for [H, W, 4], 0 for [H, W]) whose IDAT holds a zlib stream of STORED
    deflate blocks: filter byte 0 per scanline, blocks of at most 65535 bytes, Adler-32 of the raw scanlines, CRC-32 per
    chunk (PNG 1.2 / RFC 1950 / RFC 1951).  Decodes to exactly `pixels` - the image scripts/extraer_dataset.py:192,197
    writes through plt.imsave - while leaving the compression to whoever wants it."""
    import struct
    import zlib
    a = np.ascontiguousarray(pixels, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    H, W, ch = a.shape
    assert ch in (1, 4)
    raw = np.concatenate([np.zeros((H, 1), np.uint8), a.reshape(H, W * ch)], axis=1).tobytes()
    z = bytearray(b"\x78\x01")
    nblk = max(1, -(-len(raw) // 65535))
    for k in range(nblk):
        blk = raw[k * 65535:(k + 1) * 65535]
        z += struct.pack("<BHH", 1 if k == nblk - 1 else 0, len(blk), len(blk) ^ 0xFFFF) + blk
    z += struct.pack(">I", zlib.adler32(raw) & 0xFFFFFFFF)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    ihdr = struct.pack(">IIBBBBB", W, H, 8, 6 if ch == 4 else 0, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", bytes(z)) + chunk(b"IEND", b"")


# --------------------------------------------------------------------------------------
# R0  YOLO instance masks -> one predicted slice mask (the producer of R1's input)
# --------------------------------------------------------------------------------------

def resize_nearest_index(dst: int, src: int) -> np.ndarray:
    """Source index of every destination index for cv2.resize(..., interpolation=INTER_NEAREST)
    (OpenCV imgproc/resize.cpp resizeNN: sx = min(cvFloor(x * ifx), src - 1), ifx = 1 / (dst / src) in double)."""
    inv_scale = float(dst) / float(src)
    ifx = 1.0 / inv_scale
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * ifx).astype(np.int64), src - 1)


def combinar_predicciones(predicciones, shape) -> np.ndarray:
    """scripts/generar_predicciones.py:123-133: OR of the instance masks (> 0.5), each resized to `shape` =
    (height, width) with nearest-neighbour sampling.  uint8 {0, 1}, PNG orientation (the model saw the PNG)."""
    height, width = shape
    out = np.zeros((height, width), dtype=np.uint8)
    for pred in predicciones:
        pred = np.asarray(pred)
        binary = (pred > 0.5).astype(np.uint8)
        sy = resize_nearest_index(height, binary.shape[0])
        sx = resize_nearest_index(width, binary.shape[1])
        out = np.maximum(out, binary[sy][:, sx])
    return out


def normalizar_prediccion(pred: np.ndarray) -> np.ndarray:
    """scripts/generar_predicciones.py:136-140: cv2.flip(pred.T, 1) * 255 -> slice orientation, {0, 255}."""
    return (pred.T[:, ::-1] * np.uint8(255)).astype(np.uint8)


# --------------------------------------------------------------------------------------
# R1 / R2  reconstruction
# --------------------------------------------------------------------------------------

def preprocesar_mascara_pred(img_array: np.ndarray) -> np.ndarray:
    """scripts/reconstruir_volumen.py:136-150 after the PNG decode."""
    if img_array.ndim > 2:
        img_array = img_array[:, :, 0]
    if np.max(img_array) > 1:
        img_array = (img_array > 0).astype(np.float32)
    return img_array


def validar_corte(indice: int, shape_2d, shape_original, plano: str) -> None:
    """scripts/reconstruir_volumen.py:153-176."""
    max_indices = {"axial": shape_original[2], "coronal": shape_original[1], "sagital": shape_original[0]}
    if indice < 0 or indice >= max_indices[plano]:
        raise ValueError(f"Índice {indice} fuera de rango para plano {plano}.")
    expected = {
        "axial": (shape_original[0], shape_original[1]),
        "coronal": (shape_original[0], shape_original[2]),
        "sagital": (shape_original[1], shape_original[2]),
    }[plano]
    if tuple(shape_2d) != expected:
        raise ValueError(
            f"Dimensiones {tuple(shape_2d)} incorrectas para plano {plano}. Se esperaba {expected}.")


def reconstruir(slices, indices, shape_original, plano: str) -> np.ndarray:
    """scripts/reconstruir_volumen.py:199-213 without the file I/O: float32 volume (X, Y, Z);
    `slices` are decoded pred-mask arrays in slice orientation, sorted by index like :131."""
    vol = np.zeros(shape_original, dtype=np.float32)
    order = np.argsort(np.asarray(indices), kind="stable")
    for n in order:
        idx = int(indices[n])
        q = preprocesar_mascara_pred(np.asarray(slices[n]))
        validar_corte(idx, q.shape, shape_original, plano)
        if plano == "axial":
            vol[:, :, idx] = q
        elif plano == "coronal":
            vol[:, idx, :] = q
        else:
            vol[idx, :, :] = q
    return vol


# --------------------------------------------------------------------------------------
# R3  consensus
# --------------------------------------------------------------------------------------

def combinar_volumenes(axial_vol, coronal_vol, sagital_vol, umbral: int = 2) -> np.ndarray:
    """scripts/generar_consenso.py:106-109."""
    return ((axial_vol + coronal_vol + sagital_vol) >= umbral).astype(np.uint8)


# --------------------------------------------------------------------------------------
# R4  voxel confusion counts -> DSC / AUC / precision / recall
# --------------------------------------------------------------------------------------

def confusion_counts(gt: np.ndarray, pred: np.ndarray):
    """tp, fp, fn, tn as the reference's boolean sums (utils/utils.py:465-466, 474-475)."""
    tp = int(np.sum((gt == 1) & (pred == 1)))
    fp = int(np.sum((gt == 0) & (pred == 1)))
    fn = int(np.sum((gt == 1) & (pred == 0)))
    tn = int(np.sum((gt == 0) & (pred == 0)))
    return tp, fp, fn, tn


def DSC(y_true, y_pred) -> float:
    """utils/utils.py:455-460."""
    intersection = np.sum(y_true * y_pred)
    dsc = (2.0 * intersection) / (np.sum(y_true) + np.sum(y_pred) + 1e-8)
    return float(np.round(dsc, 3))


def precision(y_true, y_pred) -> float:
    """utils/utils.py:463-469."""
    tp = np.sum((y_true == 1) & (y_pred == 1))
    fp = np.sum((y_true == 0) & (y_pred == 1))
    return float(np.round(tp / (tp + fp + 1e-8), 3))


def recall(y_true, y_pred) -> float:
    """utils/utils.py:472-478."""
    tp = np.sum((y_true == 1) & (y_pred == 1))
    fn = np.sum((y_true == 1) & (y_pred == 0))
    return float(np.round(tp / (tp + fn + 1e-8), 3))


def auc_binary_from_counts(tp: int, fp: int, fn: int, tn: int) -> float:
    """roc_auc_score for a {0,1}-valued score, restated from scikit-learn's
    `_binary_clf_curve` -> `roc_curve` -> `auc` (np.trapezoid) on the three ROC points
    (0,0), (fp/(fp+tn), tp/(tp+fn)), (1,1).  (utils/utils.py:481-495; sklearn pinned 1.5.2,
    requirements.txt:37.)  Unrounded float64."""
    fps = np.array([0.0, float(fp), float(fp + tn)])
    tps = np.array([0.0, float(tp), float(tp + fn)])
    if (tp + fp) == 0 or (fn + tn) == 0:      # score has a single distinct value
        fps = fps[[0, 2]]
        tps = tps[[0, 2]]
    fpr = fps / fps[-1]
    tpr = tps / tps[-1]
    d = np.diff(fpr)
    return float((d * (tpr[1:] + tpr[:-1]) / 2.0).sum())


def AUC(y_true, y_pred) -> float:
    """utils/utils.py:481-495 for binary predictions (NaN when y_true has one class)."""
    y_true = np.asarray(y_true).ravel()
    y_pred = np.asarray(y_pred).ravel()
    if len(np.unique(y_true)) < 2:
        return float("nan")
    tp, fp, fn, tn = confusion_counts(y_true, y_pred)
    return float(np.round(auc_binary_from_counts(tp, fp, fn, tn), 3))


def metricas_desde_conteos(tp: int, fp: int, fn: int, tn: int) -> dict:
    """scripts/eval.py:115-128 evaluated from the four counts (binary volumes):
    sum(gt*p) = tp, sum(gt) = tp+fn, sum(p) = tp+fp, all exact in float64."""
    tp64, fp64, fn64 = np.int64(tp), np.int64(fp), np.int64(fn)
    inter = np.float64(tp)
    dsc = (2.0 * inter) / (np.float64(tp + fn) + np.float64(tp + fp) + 1e-8)
    prec = tp64 / (tp64 + fp64 + 1e-8)
    rec = tp64 / (tp64 + fn64 + 1e-8)
    if (tp + fn) == 0 or (fp + tn) == 0:
        auc = float("nan")
    else:
        auc = float(np.round(auc_binary_from_counts(tp, fp, fn, tn), 3))
    return {
        "DSC": float(np.round(dsc, 3)),
        "AUC": auc,
        "Precision": float(np.round(prec, 3)),
        "Recall": float(np.round(rec, 3)),
    }


def generar_diccionario_metricas(gt_vol, pred_vol) -> dict:
    """scripts/eval.py:115-128 on full arrays (the slow, literal way)."""
    return {
        "DSC": DSC(gt_vol, pred_vol),
        "AUC": AUC(gt_vol, pred_vol),
        "Precision": precision(gt_vol, pred_vol),
        "Recall": recall(gt_vol, pred_vol),
    }


# --------------------------------------------------------------------------------------
# R5 / R6 / R7  fold statistics and fold assignment
# --------------------------------------------------------------------------------------

def calcular_promedio(metricas_dic: dict) -> dict:
    """scripts/eval.py:144-160 (population std, values rounded to 3 dp)."""
    if not metricas_dic:
        raise ValueError("El diccionario de métricas está vacío.")
    return {m: {"media": float(np.round(np.mean(v), 3)), "std": float(np.round(np.std(v), 3))}
            for m, v in metricas_dic.items()}


def calcular_resumen_experimento(metricas_fold: dict) -> dict:
    """scripts/promediar_folds.py:126-134 (sample std, ddof=1)."""
    return {m: {"media": float(np.round(np.mean(v), 3)), "std": float(np.round(np.std(v, ddof=1), 3))}
            for m, v in metricas_fold.items()}


def calcular_fold(paciente_id: str, k_folds: int = 5, n_ids: int = 53) -> int:
    """utils/utils.py:299-316; `n_ids` (default 53 = reference behaviour) generalises the
    hard-coded P1..P53 range for the synthetic 75-/1024-volume cohorts."""
    numero = int(paciente_id[1:])
    folds = np.array_split(list(range(1, n_ids + 1)), k_folds)
    for i, fold in enumerate(folds, 1):
        if numero in fold:
            return i
    raise ValueError(f"No se puede calcular el fold del paciente {paciente_id}.")
