"""Minimal NIfTI-1 single-file (.nii / .nii.gz) reader and writer.

Host-side file I/O is out of scope as a kernel but the drop-in shims need
`nib.load(path).get_fdata()`, `.shape`, `.affine` and `nib.save(Nifti1Image(vol, affine), path)`
(reference utils/utils.py:153-181, utils/Paciente.py:168,179).  The shims themselves decode and encode the files on
the GPU (mslesseg_b200.codec); this host-side reader / writer (Python gzip) is what tests and tools use to make and
check such files without a GPU: float32 / uint8 / int16 / float64 data, little- or big-endian, sform or qform
affine, scl_slope / scl_inter with nibabel's rule (slope 0 or non-finite = no scaling at all).
"""
from __future__ import annotations

import gzip
import struct
from pathlib import Path

import numpy as np

_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}
_CODES = {np.dtype(v).newbyteorder("=").str[1:]: k for k, v in _DTYPES.items()}


class ImageFileError(Exception):
    """Stand-in for nibabel.filebasedimages.ImageFileError."""


def _read_all(path) -> bytes:
    path = str(path)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        return f.read()


def read_header(path):
    """(shape, affine, dtype, vox_offset, endian, slope, inter) of a NIfTI-1 file."""
    try:
        raw = _read_all(path)
    except (OSError, EOFError) as e:
        raise ImageFileError(f"Cannot read {path}: {e}") from e
    if len(raw) < 348:
        raise ImageFileError(f"{path} is not a NIfTI-1 file")
    endian = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(endian + "i", raw[:4])[0] != 348:
        raise ImageFileError(f"{path} is not a NIfTI-1 file")
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype = struct.unpack(endian + "h", raw[70:72])[0]
    pixdim = struct.unpack(endian + "8f", raw[76:108])
    vox_offset = int(struct.unpack(endian + "f", raw[108:112])[0])
    slope, inter = struct.unpack(endian + "2f", raw[112:120])
    qform_code, sform_code = struct.unpack(endian + "2h", raw[252:256])
    if datatype not in _DTYPES:
        raise ImageFileError(f"{path}: unsupported NIfTI datatype {datatype}")
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    affine = np.eye(4)
    if sform_code > 0:
        affine[0] = struct.unpack(endian + "4f", raw[280:296])
        affine[1] = struct.unpack(endian + "4f", raw[296:312])
        affine[2] = struct.unpack(endian + "4f", raw[312:328])
    elif qform_code > 0:
        b, c, d = struct.unpack(endian + "3f", raw[256:268])
        off = struct.unpack(endian + "3f", raw[268:280])
        a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
        rot = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                        [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                        [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine[:3, :3] = rot * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac])
        affine[:3, 3] = off
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return shape, affine, np.dtype(endian + _DTYPES[datatype]), vox_offset, raw, slope, inter


def load(path, dtype=None):
    """(array, affine).  The array has the file's shape in Fortran order (x fastest) - the same bytes the
    device layout [Z][Y][X] uses.  dtype=None keeps the on-disk type; np.float64 mimics get_fdata()."""
    shape, affine, dt, off, raw, slope, inter = read_header(path)
    n = int(np.prod(shape))
    arr = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    scaled = bool(np.isfinite(slope) and slope != 0.0 and np.isfinite(inter) and (slope != 1.0 or inter != 0.0))
    if scaled:
        arr = arr.astype(np.float64) * slope + inter
    if dtype is not None:
        arr = arr.astype(dtype)
    elif not scaled:
        arr = arr.astype(dt.newbyteorder("="))
    return np.asfortranarray(arr), affine


def shape_affine(path):
    """nifti.shape, nifti.affine without decoding the voxels (utils/utils.py:162-170)."""
    shape, affine, *_ = read_header(path)
    return shape, affine


def save(volumen: np.ndarray, affine, path) -> None:
    """nib.save(nib.Nifti1Image(volumen, affine), path) for float32 / uint8 / int16 / float64 arrays."""
    vol = np.asarray(volumen)
    if vol.dtype == np.bool_:
        vol = vol.astype(np.uint8)
    key = vol.dtype.newbyteorder("=").str[1:]
    if key not in _CODES:
        raise ImageFileError(f"cannot store dtype {vol.dtype} in a NIfTI-1 file")
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [vol.ndim] + list(vol.shape) + [1] * (7 - vol.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, _CODES[key])
    struct.pack_into("<h", hdr, 72, vol.dtype.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(axis=0))
    struct.pack_into("<8f", hdr, 76, 1.0, *[float(z) for z in zooms], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    hdr[123] = 2                                         # xyzt_units: mm
    struct.pack_into("<2h", hdr, 252, 0, 2)              # qform_code 0, sform_code 2 (aligned)
    for r in range(3):
        struct.pack_into("<4f", hdr, 280 + 16 * r, *[float(x) for x in affine[r]])
    hdr[344:348] = b"n+1\x00"
    payload = bytes(hdr) + b"\x00" * 4 + np.asfortranarray(vol).astype(vol.dtype.newbyteorder("<")).tobytes(order="F")
    path = str(path)
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    if path.endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=6) as f:
            f.write(payload)
    else:
        with open(path, "wb") as f:
            f.write(payload)
