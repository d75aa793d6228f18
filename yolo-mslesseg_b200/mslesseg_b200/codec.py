"""Files in, files out: the host side of the device byte-stream codec (msl_codec.cu / msl_inflate.cu).

The host only walks container headers (gzip member boundaries, PNG chunk lists, the 348-byte NIfTI header) - a few
hundred bytes of metadata per file; every payload byte is inflated / deflated, unfiltered and converted on the GPU.
There is no CPU fallback: non-conforming inputs raise.

  * .nii.gz  (reference utils/Paciente.py:168,179; utils/utils.py:153-181): `nifti_load_device`, `nifti_save_device`.
    Files written here are a sequence of gzip members of 64 KB of payload each (any gzip reader accepts them); each
    member's header carries an FEXTRA subfield 'M','S' with its compressed and raw size, so `gzip_members` finds the
    boundaries without decoding and the members are inflated in parallel, one warp each.  Foreign .gz files (one
    member, no index) are decoded by a single warp - correct, but serial.
  * PNG      (reference scripts/reconstruir_volumen.py:141; utils/utils.py:391; scripts/extraer_dataset.py:192,197):
    `png_decode_first_channel` (8-bit gray / gray+alpha / RGB / RGBA, non-interlaced) and ops.png_encode.
"""
from __future__ import annotations

import ctypes as C
import struct
import zlib
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops

CHUNK = 65536
_NIFTI_DT = {2: ("u1", 1), 4: ("i2", 2), 8: ("i4", 4), 16: ("f4", 4), 64: ("f8", 8), 256: ("i1", 1), 512: ("u2", 2), 768: ("u4", 4)}
_NIFTI_CODE = {"uint8": 2, "int16": 4, "int32": 8, "float32": 16, "float64": 64, "int8": 256, "uint16": 512, "uint32": 768}
_INF_ERR = {1: "bad container header", 2: "bad block", 3: "bad Huffman code", 4: "distance too far back", 5: "output does not fit",
            6: "input ended early"}


class CodecError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ containers (host)
def gzip_members(data: bytes) -> Optional[List[Tuple[int, int, int]]]:
    """[(offset, member bytes, raw bytes)] when `data` is a sequence of gzip members that all carry the 'MS' index
    subfield (files written by this package); None for any other gzip file."""
    out, p, n = [], 0, len(data)
    while p < n:
        if n - p < 24 or data[p:p + 4] != b"\x1f\x8b\x08\x04" or data[p + 10:p + 16] != b"\x0c\x00MS\x08\x00":
            return None
        msize, raw = struct.unpack_from("<II", data, p + 16)
        if msize < 32 or p + msize > n:
            return None
        out.append((p, msize, raw))
        p += msize
    return out or None


def png_parse(data: bytes) -> Tuple[int, int, int, bytes]:
    """(width, height, bytes per pixel, concatenated IDAT payload = one zlib stream) of an 8-bit non-interlaced PNG."""
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise CodecError("not a PNG file")
    p, idat, ihdr = 8, [], None
    while p + 8 <= len(data):
        ln, typ = struct.unpack_from(">I4s", data, p)
        body = data[p + 8:p + 8 + ln]
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        p += 12 + ln
    if ihdr is None or not idat:
        raise CodecError("PNG without IHDR / IDAT")
    w, h, depth, ctype, comp, filt, interlace = ihdr
    bpp = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if depth != 8 or bpp is None or interlace != 0 or comp != 0 or filt != 0:
        raise CodecError(f"PNG variant not supported by the device decoder (bit depth {depth}, colour type {ctype}, interlace {interlace})")
    return w, h, bpp, b"".join(idat)


# ------------------------------------------------------------------------------------------------ device calls
def _pack_streams(pieces: Sequence[bytes], device) -> Tuple[torch.Tensor, np.ndarray]:
    """Streams back to back in one pinned host buffer -> one H2D copy.  Returns (device bytes, offsets [n + 1])."""
    off = np.zeros(len(pieces) + 1, np.int64)
    np.cumsum([len(b) for b in pieces], out=off[1:])
    total = int(off[-1])
    host = torch.empty(((total + 3) & ~3) + 8, dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    hv[total:] = 0
    for i, b in enumerate(pieces):
        hv[off[i]:off[i + 1]] = np.frombuffer(b, np.uint8)
    return host.to(device, non_blocking=True), off


def inflate(pieces: Sequence[bytes], raw_sizes: Sequence[int], container: str, device) -> Tuple[torch.Tensor, np.ndarray]:
    """Inflates n streams on the device (one warp each).  raw_sizes: the exact decoded size of every stream (known from
    the container: gzip ISIZE / 'MS' subfield, PNG geometry).  Returns (device bytes, offsets [n + 1]); stream i decodes
    to out[off[i]:off[i] + raw_sizes[i]] (regions start 16-byte aligned).  Raises CodecError when a stream is damaged or
    its size does not match."""
    n = len(pieces)
    if n == 0:
        return torch.empty(0, dtype=torch.uint8, device=device), np.zeros(1, np.int64)
    src, src_off = _pack_streams(pieces, device)
    dst_off = np.zeros(n + 1, np.int64)
    np.cumsum((np.asarray(raw_sizes, np.int64) + 15) & ~15, out=dst_off[1:])
    dst = torch.empty(int(dst_off[-1]) + 16, dtype=torch.uint8, device=device)
    status = torch.empty((n, 4), dtype=torch.int32, device=device)
    cid = {"raw": L.Z_RAW, "zlib": L.Z_ZLIB, "gzip": L.Z_GZIP}[container]
    so = torch.from_numpy(src_off).to(device, non_blocking=True)
    do = torch.from_numpy(dst_off).to(device, non_blocking=True)
    L.check(L.load().msl_inflate(ops._ptr(src), src.numel(), ops._ptr(so), n, cid, ops._ptr(dst), ops._ptr(do), ops._ptr(status),
                                 ops._stream()))
    st = status.cpu().numpy().astype(np.int64) & 0xffffffff
    for i in range(n):
        if st[i, 0] != 0:
            raise CodecError(f"stream {i}: {_INF_ERR.get(int(st[i, 0]), 'error')} (deflate, container {container})")
        if st[i, 1] != raw_sizes[i]:
            raise CodecError(f"stream {i}: decoded {st[i, 1]} bytes, expected {raw_sizes[i]}")
    return dst, dst_off


def png_decode_first_channel(files: Sequence[bytes], device) -> torch.Tensor:
    """uint8 [n, H, W] on the device: channel 0 of n PNG files of one geometry (what cargar_y_preprocesar_imagen keeps,
    reference scripts/reconstruir_volumen.py:141-145).  Inflate and scanline unfiltering run on the GPU."""
    if not files:
        raise CodecError("no PNG files")
    parsed = [png_parse(f) for f in files]
    w, h, bpp = parsed[0][:3]
    for q in parsed:
        if q[:3] != (w, h, bpp):
            raise CodecError(f"PNG files of different geometry: {(w, h, bpp)} vs {q[:3]}")
    raw = h * (w * bpp + 1)
    dst, dst_off = inflate([q[3] for q in parsed], [raw] * len(files), "zlib", device)
    out = torch.empty((len(files), h, w), dtype=torch.uint8, device=device)
    status = torch.empty(len(files), dtype=torch.int32, device=device)
    ro = torch.from_numpy(dst_off).to(device)
    L.check(L.load().msl_png_unfilter(ops._ptr(dst), ops._ptr(ro), len(files), h, w, bpp, ops._ptr(out), ops._ptr(status), ops._stream()))
    if int(status.max().item()) != 0:
        raise CodecError("PNG with an unknown scanline filter type")
    return out


# ------------------------------------------------------------------------------------------------ NIfTI
def _nifti_header(first: bytes):
    if len(first) < 348 or struct.unpack_from("<i", first, 0)[0] != 348:
        raise CodecError("not a little-endian NIfTI-1 file")
    dim = struct.unpack_from("<8h", first, 40)
    datatype = struct.unpack_from("<h", first, 70)[0]
    pixdim = struct.unpack_from("<8f", first, 76)
    vox_offset = int(struct.unpack_from("<f", first, 108)[0])
    slope, inter = struct.unpack_from("<2f", first, 112)
    qform_code, sform_code = struct.unpack_from("<2h", first, 252)
    if datatype not in _NIFTI_DT:
        raise CodecError(f"unsupported NIfTI datatype {datatype}")
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    affine = np.eye(4)
    if sform_code > 0:
        for r in range(3):
            affine[r] = struct.unpack_from("<4f", first, 280 + 16 * r)
    elif qform_code > 0:
        b, c, d = struct.unpack_from("<3f", first, 256)
        off = struct.unpack_from("<3f", first, 268)
        a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
        rot = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                        [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                        [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine[:3, :3] = rot * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac])
        affine[:3, 3] = off
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    scaled = bool(np.isfinite(slope) and slope != 0.0 and (slope != 1.0 or inter != 0.0))     # nibabel: slope 0 / nan = none
    return shape, affine, datatype, vox_offset, float(slope), float(inter), scaled


def nifti_read_header(path):
    """(shape, affine) from the first bytes of a .nii / .nii.gz (host: a 348-byte header, no voxel is decoded)."""
    path = str(path)
    with open(path, "rb") as f:
        head = f.read(4096)
    if path.endswith(".gz"):
        head = zlib.decompressobj(wbits=31).decompress(head, 348)
    shape, affine, *_ = _nifti_header(head)
    return shape, affine


def nifti_load_device(path, device, dtype=torch.float32):
    """(volume [Z][Y][X] on the device as float32, uint8 or float64, shape (X, Y, Z), affine).  The file bytes are uploaded
    as they are; inflate + datatype conversion (+ scl_slope / scl_inter) run on the GPU.  uint8 output demands integral
    values in 0..255 (masks); float32 output demands float32-representable values - otherwise CodecError; float64 holds
    whatever nib.load(path).get_fdata() yields."""
    path = str(path)
    data = Path(path).read_bytes()
    dev = torch.device(device)
    if path.endswith(".gz"):
        members = gzip_members(data)
        if members is not None:
            pieces = [data[o:o + m] for o, m, _ in members]
            sizes = [r for _, _, r in members]
        else:
            if len(data) < 18:
                raise CodecError(f"{path}: not a gzip file")
            pieces, sizes = [data], [struct.unpack_from("<I", data, len(data) - 4)[0]]
        head = zlib.decompressobj(wbits=31).decompress(data[:4096], 348)
        raw, off = inflate(pieces, sizes, "gzip", dev)
        contiguous = all(s % 16 == 0 for s in sizes[:-1])
        if not contiguous:       # members whose size is not a multiple of 16 leave gaps: close them
            parts = [raw[int(off[i]):int(off[i]) + sizes[i]] for i in range(len(sizes))]
            raw = torch.cat(parts)
    else:
        head = data[:348]
        raw = torch.from_numpy(np.frombuffer(data, np.uint8).copy()).to(dev)
    shape, affine, datatype, vox_offset, slope, inter, scaled = _nifti_header(head)
    if len(shape) != 3:
        raise CodecError(f"{path}: expected a 3-D volume, got shape {shape}")
    nvox = int(np.prod(shape))
    if raw.numel() < vox_offset + nvox * _NIFTI_DT[datatype][1]:
        raise CodecError(f"{path}: file shorter than its header announces")
    X, Y, Z = shape
    out = torch.empty((Z, Y, X), dtype=dtype, device=dev)
    inexact = torch.zeros(1, dtype=torch.int64, device=dev)
    payload = raw[vox_offset:]
    L.check(L.load().msl_nifti_convert(ops._ptr(payload), datatype, nvox, slope, inter, 1 if scaled else 0,
                                       ops._ptr(out) if dtype == torch.float32 else None,
                                       ops._ptr(out) if dtype == torch.uint8 else None,
                                       ops._ptr(out) if dtype == torch.float64 else None, ops._ptr(inexact), ops._stream()))
    bad = int(inexact.item())
    if bad:
        raise CodecError(f"{path}: {bad} voxels are not representable as {dtype}")
    return out, shape, affine


def nifti_header_bytes(shape_xyz, np_dtype, affine) -> bytes:
    """352 bytes: NIfTI-1 header + 4-byte extension flag, as nib.Nifti1Image(vol, affine) would describe the volume."""
    dt = np.dtype(np_dtype)
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, len(shape_xyz), *shape_xyz, *([1] * (7 - len(shape_xyz))))
    struct.pack_into("<2h", hdr, 70, _NIFTI_CODE[dt.name], dt.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(axis=0))
    struct.pack_into("<8f", hdr, 76, 1.0, *[float(z) for z in zooms], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    hdr[123] = 2
    struct.pack_into("<2h", hdr, 252, 0, 2)
    for r in range(3):
        struct.pack_into("<4f", hdr, 280 + 16 * r, *[float(x) for x in affine[r]])
    hdr[344:348] = b"n+1\x00"
    return bytes(hdr)


def nifti_gz_device(vol_zyx: torch.Tensor, affine, dist2: Optional[int] = None) -> ops.PackedStreams:
    """The bytes of a .nii.gz file for a device volume [Z][Y][X] (float32, uint8, float64, int16, int32, int8): header + voxels deflated on the GPU in
    64 KB members (reference utils/utils.py:173-181 guardar_volumen).  `b"".join(ps.files())` / ps.to_host() is the file."""
    ops._need_cuda(vol_zyx, "vol_zyx")
    Z, Y, X = (int(d) for d in vol_zyx.shape)
    npdt = {torch.float32: np.float32, torch.uint8: np.uint8, torch.float64: np.float64, torch.int16: np.int16,
            torch.int32: np.int32, torch.int8: np.int8}[vol_zyx.dtype]
    hdr = torch.from_numpy(np.frombuffer(nifti_header_bytes((X, Y, Z), npdt, affine), np.uint8).copy()).to(vol_zyx.device)
    buf = torch.cat([hdr, vol_zyx.contiguous().reshape(-1).view(torch.uint8)])
    if dist2 is None:
        dist2 = vol_zyx.element_size() if vol_zyx.element_size() > 1 else 0
    return ops.deflate_chunks(buf, chunk_len=CHUNK, container="gzip", dist2=dist2)


def nifti_save_device(vol_zyx: torch.Tensor, affine, path) -> int:
    """Writes the .nii.gz; returns the file size."""
    data, off = nifti_gz_device(vol_zyx, affine).to_host()
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(data[:int(off[-1])].tobytes())
    return int(off[-1])
