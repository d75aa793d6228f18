"""TEST INFRASTRUCTURE - not product code.

Stand-ins for the three third-party packages the reference imports at module top level and that are absent
from this image (no network, no wheel): nibabel 5.3.2, matplotlib 3.10.0, ultralytics 8.3.70
(requirements.txt:23,20,47).  `install(functional=True)` registers them in `sys.modules` so that the UNMODIFIED
reference stage scripts run end to end on a synthetic dataset tree:

* nibabel       -> NIfTI-1 single-file reader / writer (load().get_fdata() / .shape / .affine, Nifti1Image,
                   save, filebasedimages.ImageFileError).  File format only - pinned by the demo volumes.
* matplotlib    -> pyplot.imsave(path, A, cmap="gray", origin="lower"): oracle.imsave restatement (E8,
                   **parity unpinned**: restated from matplotlib 3.10's Normalize / Colormap), written with Pillow.
* ultralytics   -> data.converter.convert_segment_masks_to_yolo_seg restated from ultralytics 8.3.70
                   (cv2.findContours RETR_EXTERNAL / CHAIN_APPROX_SIMPLE, >= 3 points, coordinates / (w, h)
                   rounded to 6 decimals, one line per contour, pixel value v -> class v - 1).  The contour
                   extraction itself is the installed cv2; only the text formatting is restated.
"""
from __future__ import annotations

import gzip
import logging
import struct
import sys
import types
from pathlib import Path

import numpy as np

_NIFTI_DT = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}
_NIFTI_CODE = {"u1": 2, "i2": 4, "i4": 8, "f4": 16, "f8": 64, "i1": 256, "u2": 512, "u4": 768}


class ImageFileError(Exception):
    pass


class _Nifti1Image:
    def __init__(self, dataobj, affine, header=None):
        self._data = np.asanyarray(dataobj)
        self.affine = None if affine is None else np.asarray(affine, dtype=np.float64)
        self.header = header
        self.shape = tuple(self._data.shape)

    def get_fdata(self, dtype=np.float64):
        return np.asfortranarray(self._data.astype(dtype))

    @property
    def dataobj(self):
        return self._data


def _nib_load(path):
    path = str(path)
    try:
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rb") as f:
            raw = f.read()
    except FileNotFoundError:
        raise
    except (OSError, EOFError) as e:
        raise ImageFileError(f"Cannot work out file type of \"{path}\"") from e
    if len(raw) < 352 or struct.unpack("<i", raw[:4])[0] not in (348, 1543569408):
        raise ImageFileError(f"Cannot work out file type of \"{path}\"")
    en = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    dim = struct.unpack(en + "8h", raw[40:56])
    datatype = struct.unpack(en + "h", raw[70:72])[0]
    vox_offset = int(struct.unpack(en + "f", raw[108:112])[0])
    slope, inter = struct.unpack(en + "2f", raw[112:120])
    sform_code = struct.unpack(en + "h", raw[254:256])[0]
    pixdim = struct.unpack(en + "8f", raw[76:108])
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    arr = np.frombuffer(raw, dtype=np.dtype(en + _NIFTI_DT[datatype]), count=int(np.prod(shape)), offset=vox_offset)
    arr = arr.reshape(shape, order="F")
    if np.isfinite(slope) and slope != 0.0 and (slope != 1.0 or inter != 0.0):       # nibabel: slope 0 / nan = no scaling
        arr = arr.astype(np.float64) * slope + inter
    affine = np.eye(4)
    if sform_code > 0:
        for r in range(3):
            affine[r] = struct.unpack(en + "4f", raw[280 + 16 * r:296 + 16 * r])
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return _Nifti1Image(arr, affine, header=raw[:348])


def _nib_save(img, path):
    vol = np.asarray(img.dataobj)
    if vol.dtype == np.bool_:
        vol = vol.astype(np.uint8)
    key = vol.dtype.newbyteorder("=").str[1:]
    affine = np.eye(4) if img.affine is None else img.affine
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, vol.ndim, *vol.shape, *([1] * (7 - vol.ndim)))
    struct.pack_into("<2h", hdr, 70, _NIFTI_CODE[key], vol.dtype.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(axis=0))
    struct.pack_into("<8f", hdr, 76, 1.0, *[float(z) for z in zooms], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    hdr[123] = 2
    struct.pack_into("<2h", hdr, 252, 0, 2)
    for r in range(3):
        struct.pack_into("<4f", hdr, 280 + 16 * r, *[float(x) for x in affine[r]])
    hdr[344:348] = b"n+1\x00"
    payload = bytes(hdr) + b"\x00" * 4 + np.asfortranarray(vol).astype(vol.dtype.newbyteorder("<")).tobytes(order="F")
    path = str(path)
    if path.endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(payload)
    else:
        with open(path, "wb") as f:
            f.write(payload)


def _imsave(fname, arr, vmin=None, vmax=None, cmap=None, format=None, origin=None, dpi=100, **kw):
    """matplotlib.pyplot.imsave for the one way the reference calls it (extraer_dataset.py:192,197)."""
    from PIL import Image
    from oracle import oracle as O
    if cmap != "gray" or vmin is not None or vmax is not None:
        raise NotImplementedError("imsave stand-in: only cmap='gray' without vmin / vmax (reference call sites)")
    A = np.asarray(arr)
    if origin == "lower":
        A = A[::-1]
    g = O.gray_cmap_bytes()[O.imsave_gray_index(A)]
    rgba = np.empty(g.shape + (4,), np.uint8)
    rgba[..., 0] = rgba[..., 1] = rgba[..., 2] = g
    rgba[..., 3] = 255
    Image.fromarray(rgba, "RGBA").save(str(fname), format="PNG")


def yolo_seg_lines(mask: np.ndarray, classes: int) -> list:
    """The label lines ultralytics 8.3.70 convert_segment_masks_to_yolo_seg writes for one grayscale mask."""
    import cv2
    h, w = mask.shape
    pixel_to_class = {i + 1: i for i in range(classes)}
    lines = []
    for value in np.unique(mask):
        if value == 0:
            continue
        cls = pixel_to_class.get(int(value), -1)
        if cls == -1:
            continue
        contours, _ = cv2.findContours((mask == value).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for contour in contours:
            if len(contour) >= 3:
                pts = contour.squeeze()
                item = [cls]
                for point in pts:
                    item.append(round(point[0] / w, 6))
                    item.append(round(point[1] / h, 6))
                lines.append(" ".join(map(str, item)))
    return lines


def _convert_segment_masks_to_yolo_seg(masks_dir, output_dir, classes):
    import cv2
    for mask_path in Path(masks_dir).iterdir():
        if mask_path.suffix in {".png", ".jpg"}:
            mask = cv2.imread(str(mask_path), cv2.IMREAD_GRAYSCALE)
            lines = yolo_seg_lines(mask, classes)
            with open(Path(output_dir) / f"{mask_path.stem}.txt", "w") as f:
                for ln in lines:
                    f.write(ln + "\n")


def install(functional: bool = True) -> None:
    """Register the stand-ins (idempotent).  functional=False registers import-only placeholders."""
    names = ["nibabel", "nibabel.filebasedimages", "ultralytics", "ultralytics.utils", "ultralytics.data",
             "ultralytics.data.converter", "matplotlib", "matplotlib.pyplot"]
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    nib, fb = sys.modules["nibabel"], sys.modules["nibabel.filebasedimages"]
    nib.filebasedimages = fb
    if not hasattr(fb, "ImageFileError"):
        fb.ImageFileError = ImageFileError
    ul = sys.modules["ultralytics"]
    ul.YOLO = getattr(ul, "YOLO", object)
    ul.utils, ul.data = sys.modules["ultralytics.utils"], sys.modules["ultralytics.data"]
    ul.data.converter = sys.modules["ultralytics.data.converter"]
    ul.utils.LOGGER = logging.getLogger("ultralytics-stub")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if functional:
        nib.load, nib.save, nib.Nifti1Image = _nib_load, _nib_save, _Nifti1Image
        ul.data.converter.convert_segment_masks_to_yolo_seg = _convert_segment_masks_to_yolo_seg
        sys.modules["matplotlib.pyplot"].imsave = _imsave
    elif not hasattr(ul.data.converter, "convert_segment_masks_to_yolo_seg"):
        ul.data.converter.convert_segment_masks_to_yolo_seg = lambda **k: None
