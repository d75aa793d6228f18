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

CHUNK = 16384          # payload bytes per gzip member of the files written here (one deflate tile, one decoder lane)
_NIFTI_DT = {2: ("u1", 1), 4: ("i2", 2), 8: ("i4", 4), 16: ("f4", 4), 64: ("f8", 8), 256: ("i1", 1), 512: ("u2", 2), 768: ("u4", 4)}
_NIFTI_CODE = {"uint8": 2, "int16": 4, "int32": 8, "float32": 16, "float64": 64, "int8": 256, "uint16": 512, "uint32": 768}
_INF_ERR = {1: "bad container header", 2: "bad block", 3: "bad Huffman code", 4: "distance too far back", 5: "output does not fit",
            6: "input ended early"}


class CodecError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ containers (host)
def gzip_index_member(member_sizes, raw_sizes) -> bytes:
    """A gzip member with an EMPTY payload whose FEXTRA field ('M','I') lists the compressed and raw size of every
    preceding member followed by their count: appended to a file written by this package it lets a reader find all
    member boundaries with one look at the file's tail.  Any gzip reader decodes it to zero bytes."""
    n = len(member_sizes)
    body = np.stack([np.asarray(member_sizes, "<u4"), np.asarray(raw_sizes, "<u4")], axis=1).tobytes() + struct.pack("<I", n)
    if len(body) + 4 > 65535:
        return b""                                             # too many members for one extra field: readers fall back to hopping
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", len(body) + 4) + b"MI" + struct.pack("<H", len(body)) + body
            + b"\x03\x00" + b"\x00" * 8)


def gzip_member_table(data) -> Optional[np.ndarray]:
    """int64 [n, 3] = (offset, member bytes, raw bytes) of the payload members of a file written by this package, from the
    index member at its tail (vectorised: no per-member work); None when the file has no such index."""
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
    L_ = buf.size
    if L_ < 40:
        return None
    n = int(buf[L_ - 14:L_ - 10].view("<u4")[0])               # ... sizes[n][2], n | 03 00 | crc, isize
    start = L_ - (10 + 2 + 4 + 8 * n + 4 + 2 + 8)
    if n <= 0 or start < 0 or bytes(buf[start:start + 4]) != b"\x1f\x8b\x08\x04" or bytes(buf[start + 12:start + 14]) != b"MI":
        return None
    t = buf[start + 16:start + 16 + 8 * n].view("<u4").reshape(n, 2).astype(np.int64)
    off = np.zeros(n, np.int64)
    np.cumsum(t[:-1, 0], out=off[1:])
    if off[-1] + t[-1, 0] != start:
        return None
    return np.concatenate([off[:, None], t], axis=1)


def gzip_members(data: bytes) -> Optional[List[Tuple[int, int, int]]]:
    """[(offset, member bytes, raw bytes)] when `data` is a sequence of gzip members that all carry the 'MS' index
    subfield (files written by this package); None for any other gzip file."""
    tab = gzip_member_table(data)
    if tab is not None:
        return [tuple(int(v) for v in row) for row in tab]
    out, p, n = [], 0, len(data)
    while p < n:
        if n - p >= 14 and data[p + 12:p + 14] == b"MI" and out:      # the index member at the tail
            break
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


@ops._nvtx("inflate")
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


# ---- asynchronous, batch-level forms (device tensors in, device tensors out, nothing synchronises): the building blocks
# of a pipelined cohort step; the status tensors are checked by the caller when the step's results are collected
@ops._nvtx("inflate_device")
def inflate_device(src: torch.Tensor, src_off: torch.Tensor, dst: torch.Tensor, dst_off: torch.Tensor, container: str,
                   status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """msl_inflate on device tensors: stream i = src[src_off[i]:src_off[i+1]] -> dst[dst_off[i]:dst_off[i+1]] (int64 offset
    tensors of n + 1 entries on the device).  Returns the status tensor int32 [n, 4] (see include/mslesseg.h)."""
    n = int(src_off.numel()) - 1
    if status is None:
        status = torch.empty((n, 4), dtype=torch.int32, device=src.device)
    cid = {"raw": L.Z_RAW, "zlib": L.Z_ZLIB, "gzip": L.Z_GZIP}[container]
    L.check(L.load().msl_inflate(ops._ptr(src), src.numel(), ops._ptr(src_off), n, cid, ops._ptr(dst), ops._ptr(dst_off),
                                 ops._ptr(status), ops._stream()))
    return status


@ops._nvtx("png_unfilter_device")
def png_unfilter_device(raw: torch.Tensor, raw_off: torch.Tensor, H: int, W: int, bpp: int, out: torch.Tensor,
                        status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """msl_png_unfilter on device tensors: n = out.shape[0] inflated images -> out uint8 [n, H, W] (first channel)."""
    n = int(out.shape[0])
    if status is None:
        status = torch.empty(n, dtype=torch.int32, device=raw.device)
    L.check(L.load().msl_png_unfilter(ops._ptr(raw), ops._ptr(raw_off), n, H, W, bpp, ops._ptr(out), ops._ptr(status), ops._stream()))
    return status


@ops._nvtx("nifti_convert_device")
def nifti_convert_device(payload: torch.Tensor, datatype: int, out: torch.Tensor, inexact: torch.Tensor, slope: float = 1.0,
                         inter: float = 0.0, scaled: bool = False) -> None:
    """msl_nifti_convert: out.numel() voxels of NIfTI datatype `datatype` at payload -> out (float32 / uint8 / float64)."""
    L.check(L.load().msl_nifti_convert(ops._ptr(payload), int(datatype), out.numel(), float(slope), float(inter), 1 if scaled else 0,
                                       ops._ptr(out) if out.dtype == torch.float32 else None,
                                       ops._ptr(out) if out.dtype == torch.uint8 else None,
                                       ops._ptr(out) if out.dtype == torch.float64 else None, ops._ptr(inexact), ops._stream()))


def png_table(buf: np.ndarray, file_off: np.ndarray):
    """Vectorised look at n PNG files stored back to back in `buf` (file i = buf[file_off[i]:file_off[i+1]]) that share
    the simplest layout - signature, IHDR, ONE IDAT, IEND (what cv2.imwrite and ops.png_encode produce for small images).
    Returns (width, height, bytes per pixel, idat_start int64 [n], idat_len int64 [n]) or None when a file deviates (the
    caller then parses file by file with png_parse)."""
    starts = np.asarray(file_off[:-1], np.int64)
    sizes = np.diff(np.asarray(file_off, np.int64))
    if starts.size == 0 or (sizes < 57).any():
        return None
    head = buf[starts[:, None] + np.arange(41)[None, :]]                      # [n, 41]
    sig = np.frombuffer(b"\x89PNG\r\n\x1a\n\x00\x00\x00\rIHDR", np.uint8)
    if not (head[:, :16] == sig).all() or not (head[:, 37:41] == np.frombuffer(b"IDAT", np.uint8)).all():
        return None
    be = lambda a: (a[:, 0].astype(np.int64) << 24) | (a[:, 1].astype(np.int64) << 16) | (a[:, 2].astype(np.int64) << 8) | a[:, 3]
    w, h, ilen = be(head[:, 16:20]), be(head[:, 20:24]), be(head[:, 33:37])
    if not ((w == w[0]).all() and (h == h[0]).all() and (head[:, 24:29] == head[0, 24:29]).all() and (ilen + 57 == sizes).all()):
        return None
    depth, ctype, comp, filt, interlace = (int(v) for v in head[0, 24:29])
    bpp = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if depth != 8 or bpp is None or interlace or comp or filt:
        raise CodecError(f"PNG variant not supported by the device decoder (bit depth {depth}, colour type {ctype}, interlace {interlace})")
    return int(w[0]), int(h[0]), bpp, starts + 41, ilen


@ops._nvtx("png_decode_first_channel")
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


@ops._nvtx("nifti_load_device")
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


@ops._nvtx("nifti_gz_device")
def nifti_gz_device(vol: torch.Tensor, affine, dist2: Optional[int] = None, como_float32: bool = False,
                    out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> ops.PackedStreams:
    """The gzip members of the .nii.gz file(s) of device volumes [Z][Y][X] or [n][Z][Y][X] (float32, uint8, float64, int16,
    int32, int8): header (a 352-byte prefix, never concatenated with the voxels) + voxels deflated on the GPU in 64 KB
    members (reference utils/utils.py:173-181 guardar_volumen).  como_float32: the volume is a uint8 {0, 1} mask but the
    file stores float32 (what reconstruir_volumen writes).  ps.streams_per_file members belong to each volume."""
    ops._need_cuda(vol, "vol")
    v4 = vol if vol.dim() == 4 else vol[None]
    n, Z, Y, X = (int(d) for d in v4.shape)
    npdt = {torch.float32: np.float32, torch.uint8: np.uint8, torch.float64: np.float64, torch.int16: np.int16,
            torch.int32: np.int32, torch.int8: np.int8}[v4.dtype]
    if como_float32:
        if v4.dtype != torch.uint8:
            raise ValueError("como_float32 expects a uint8 mask volume")
        npdt = np.float32
    hdr = torch.from_numpy(np.frombuffer(nifti_header_bytes((X, Y, Z), npdt, affine), np.uint8).copy()).to(v4.device, non_blocking=True)
    if dist2 is None:
        dist2 = np.dtype(npdt).itemsize if np.dtype(npdt).itemsize > 1 else 0
    return ops.deflate_files(v4.contiguous(), prefix=hdr, expand_u8_to_f32=como_float32, chunk_len=CHUNK, container="gzip",
                             dist2=dist2, out=out, workspace=workspace)


def nifti_gz_bytes(ps: ops.PackedStreams, archivo: int = 0) -> bytes:
    """The complete .nii.gz of file `archivo` of nifti_gz_device: its members + the index member (gzip_index_member)."""
    data, off = ps.to_host()
    meta = ps.meta.cpu().numpy().astype(np.int64) & 0xffffffff
    spv = getattr(ps, "streams_per_file", len(off) - 1)
    a, b = archivo * spv, (archivo + 1) * spv
    return data[int(off[a]):int(off[b])].tobytes() + gzip_index_member(np.diff(off[a:b + 1]), meta[a:b, 1])


def nifti_save_device(vol_zyx: torch.Tensor, affine, path) -> int:
    """Writes the .nii.gz; returns the file size."""
    blob = nifti_gz_bytes(nifti_gz_device(vol_zyx, affine))
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(blob)
    return len(blob)


# ------------------------------------------------------------------------------------------------ dataset preparation
def reindex_gz(path, dest=None, device=None) -> bool:
    """Rewrites a .gz file from another writer (one member, one long deflate stream: a single warp's serial work, about a
    second per MSLesSeg volume) as the member-indexed file this package writes (16 KB of raw bytes per member + the index
    member: one warp per member, a fraction of a millisecond per volume).  Both the decode and the encode run on the GPU; the
    decoded bytes are unchanged, so nibabel / gzip read the new file exactly like the old one.  Returns False when the file
    already carries the index.  Meant to be run once over a dataset (`python -m mslesseg_b200.codec reindex <files>`)."""
    path = Path(path)
    blob = path.read_bytes()
    if gzip_member_table(blob) is not None:
        return False
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    members = gzip_members(blob)                  # walks the members when they carry their size, else None
    if members is None:
        isize = int.from_bytes(blob[-4:], "little")      # ISIZE of the last member = the file for a single-member file < 4 GB
        dst, _ = inflate([blob], [isize], "gzip", device)
        raw = dst[:isize]
    else:
        pieces = [blob[o:o + m] for o, m, _ in members]
        dst, off = inflate(pieces, [r for _, _, r in members], "gzip", device)
        raw = torch.cat([dst[int(off[i]):int(off[i]) + members[i][2]] for i in range(len(members))])
    elem = 1
    if raw.numel() >= 352 and int.from_bytes(raw[:4].cpu().numpy().tobytes(), "little") == 348:      # NIfTI-1: bitpix at byte 72
        elem = max(1, min(4, int.from_bytes(raw[72:74].cpu().numpy().tobytes(), "little") // 8))
    ps = ops.deflate_chunks(raw.contiguous(), chunk_len=CHUNK, container="gzip", dist2=elem)
    out = nifti_gz_bytes(ps)
    dest = Path(dest) if dest is not None else path
    tmp = dest.with_name(dest.name + ".tmp")
    tmp.write_bytes(out)
    tmp.replace(dest)
    return True


if __name__ == "__main__":
    import sys
    if len(sys.argv) < 3 or sys.argv[1] != "reindex":
        raise SystemExit("usage: python -m mslesseg_b200.codec reindex <file.nii.gz | directory> ...")
    todo = []
    for a in sys.argv[2:]:
        todo += sorted(Path(a).rglob("*.nii.gz")) if Path(a).is_dir() else [Path(a)]
    for f in todo:
        print(("reindexed " if reindex_gz(f) else "kept      ") + str(f))
