"""Batched device-tensor API over the C ABI (include/mslesseg.h).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic step runs in
the hand-written sm_100a kernels of libmslesseg.so.  All functions enqueue on
`torch.cuda.current_stream()` and return device tensors without synchronising.

Layouts: volumes are [nvol][Z][Y][X] (x fastest) - the same bytes as the reference's
Fortran-ordered (X, Y, Z) arrays.  Slice stacks are [n][rows][cols] in slice orientation
("G") or [n][cols][rows] in PNG orientation ("P", P[r, c] = G[c, cols-1-r]).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import tables as T

PLANOS = ("axial", "coronal", "sagital")
MEJORAS = ("HE", "CLAHE", "GC", "LT")
_LAYOUT_ID = {"G": L.OUT_G, "P": L.OUT_P, "PNG_GRAY": L.OUT_PNG_GRAY, "PNG_RGBA": L.OUT_PNG_RGBA}
_tables_cache: Dict[tuple, torch.Tensor] = {}


def _nvtx(name: str):
    """Decorator: an NVTX range per stage call, so that an nsys / ncu timeline of a cohort step reads as
    stages.  torch.cuda.nvtx is a no-op without a profiler attached."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapper(*a, **k):
            torch.cuda.nvtx.range_push("msl." + name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapper
    return deco


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _dtype_id(t: torch.Tensor, name: str) -> int:
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.uint8:
        return L.U8
    raise TypeError(f"{name} must be float32 or uint8, got {t.dtype}")


def device_tables(device, lut_out: str = "gray") -> torch.Tensor:
    """The MSL_TABLES_BYTES constant block, uploaded once per device (and per CLAHE output table, see
    tables.host_tables)."""
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), lut_out)
    t = _tables_cache.get(key)
    if t is None:
        t = torch.from_numpy(T.host_tables(lut_out).copy()).to(dev)
        _tables_cache[key] = t
    return t


def plane_dims(plano: str, X: int, Y: int, Z: int) -> Tuple[int, int, int]:
    """(n_slices, rows, cols) of a plane (reference utils/Paciente.py:186-193, :240-244)."""
    if plano == "axial":
        return Z, X, Y
    if plano == "coronal":
        return Y, X, Z
    if plano == "sagital":
        return X, Y, Z
    raise ValueError(f"Plano {plano} no válido.")


def _index_tensor(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.int32).contiguous()
    return torch.as_tensor(np.asarray(a, dtype=np.int32), device=device)


def _out_shape(n: int, rows: int, cols: int, layout: str):
    if layout == "G":
        return (n, rows, cols)
    if layout in ("P", "PNG_GRAY"):
        return (n, cols, rows)
    if layout == "PNG_RGBA":
        return (n, cols, rows, 4)
    raise ValueError(f"layout {layout!r} not in {sorted(_LAYOUT_ID)}")


# ------------------------------------------------------------------------------------ E0
@_nvtx("lesion_slices")
def lesion_slices(gt: torch.Tensor):
    """any(mask_slice > 0) for every slice of the three planes.  gt: [nvol, Z, Y, X] uint8/float32.
    Returns (any_ax [nvol, Z], any_co [nvol, Y], any_sa [nvol, X]) uint8."""
    _need_cuda(gt, "gt")
    if gt.dim() != 4:
        raise ValueError("gt must be [nvol, Z, Y, X]")
    nvol, Z, Y, X = gt.shape
    ax = torch.empty((nvol, Z), dtype=torch.uint8, device=gt.device)
    co = torch.empty((nvol, Y), dtype=torch.uint8, device=gt.device)
    sa = torch.empty((nvol, X), dtype=torch.uint8, device=gt.device)
    L.check(L.load().msl_lesion_slices(_ptr(gt), _dtype_id(gt, "gt"), nvol, X, Y, Z, _ptr(ax), _ptr(co), _ptr(sa), _stream()))
    return ax, co, sa


# ------------------------------------------------------------------------------------ E1-E8
@_nvtx("slice_ranges")
def slice_ranges(vol: torch.Tensor) -> Dict[str, torch.Tensor]:
    """{plano: float32 [nvol, n_plane, 2]} = (min, max) of every slice of the three planes in one pass over the float32
    volumes [nvol, Z, Y, X] (the statistics of normalizar_a_uint8, reference utils/utils.py:400-405)."""
    _need_cuda(vol, "vol")
    if vol.dtype != torch.float32 or vol.dim() != 4:
        raise ValueError("vol must be float32 [nvol, Z, Y, X]")
    nvol, Z, Y, X = (int(d) for d in vol.shape)
    out = torch.empty((nvol, Z + Y + X, 2), dtype=torch.float32, device=vol.device)
    L.check(L.load().msl_slice_ranges(_ptr(vol), nvol, X, Y, Z, _ptr(out), _stream()))
    return {"axial": out[:, :Z], "coronal": out[:, Z:Z + Y], "sagital": out[:, Z + Y:]}


_STACK_MIN_SLICES = 64      # slice lists at least this long go through the staged stack + dense kernel (enhance_slices)


@_nvtx("enhance_slices")
def enhance_slices(vol: torch.Tensor, mejora: Optional[str], plano: str, vol_of_slice=None, idx_of_slice=None,
                   layout: str = "G", out: Optional[torch.Tensor] = None, lut_out: str = "gray") -> torch.Tensor:
    """Enhanced slices of resident volumes.  vol: [nvol, Z, Y, X] float32 (normalised per slice like
    normalizar_a_uint8) or uint8 (used as is).  With no index lists every slice of every volume is
    produced (s = v * n_plane + i)."""
    _need_cuda(vol, "vol")
    if vol.dim() != 4:
        raise ValueError("vol must be [nvol, Z, Y, X]")
    if mejora not in L.MEJORA_ID:
        raise ValueError(f"Mejora no reconocida: {mejora}.")
    nvol, Z, Y, X = vol.shape
    n_p, rows, cols = plane_dims(plano, X, Y, Z)
    if (vol_of_slice is None) != (idx_of_slice is None):
        raise ValueError("vol_of_slice and idx_of_slice go together")
    if vol_of_slice is None:
        ns, vs, ix = nvol * n_p, None, None
    else:
        checked = False
        if not isinstance(vol_of_slice, torch.Tensor) and not isinstance(idx_of_slice, torch.Tensor):
            # host lists are range-checked here; the reference raises IndexError for such a slice index (the kernel skips
            # (volume, index) pairs outside the volumes)
            hv, hi = np.asarray(vol_of_slice, dtype=np.int64).reshape(-1), np.asarray(idx_of_slice, dtype=np.int64).reshape(-1)
            if hv.size and (hv.min() < 0 or hv.max() >= nvol):
                raise IndexError(f"volume index outside [0, {nvol})")
            if hi.size and (hi.min() < 0 or hi.max() >= n_p):
                raise IndexError(f"index {int(hi.max() if hi.max() >= n_p else hi.min())} is out of bounds for plano {plano} with size {n_p}")
            checked = True
        vs, ix = _index_tensor(vol_of_slice, vol.device), _index_tensor(idx_of_slice, vol.device)
        if vs.shape != ix.shape or vs.dim() != 1:
            raise ValueError("vol_of_slice / idx_of_slice must be 1-D and of equal length")
        ns = int(vs.numel())
    shape = _out_shape(ns, rows, cols, layout)
    if out is None:
        # index lists that live on the device cannot be checked without a synchronisation: slices whose pair lies outside
        # the volumes are skipped by the kernel and come back as zeros instead of uninitialised memory
        unchecked = vol_of_slice is not None and not checked
        out = (torch.zeros if unchecked else torch.empty)(shape, dtype=torch.uint8, device=vol.device)
    else:
        _need_cuda(out, "out")
        if tuple(out.shape) != shape or out.dtype != torch.uint8:
            raise ValueError(f"out must be uint8 {shape}")
    if ns == 0:                      # an empty tensor has a NULL data pointer, which the ABI reads as "dense"
        return out
    lib = L.load()
    tabs = device_tables(vol.device, lut_out)
    if (mejora is not None and layout == "P" and ns >= _STACK_MIN_SLICES and (rows * cols) % 4 == 0 and out.is_contiguous()
            and vol.dtype == torch.float32 and (vol_of_slice is None or checked)):
        # (uint8 volumes skip E1, so a slice's maximum need not be 255 as the dense kernel's LT row assumes; unchecked device
        # index lists keep the kernel that skips bad pairs)
        # long lists: E1 per slice into a staged stack (PNG orientation, 16-byte pitch), then the whole-volume kernel over the
        # stack - the same bytes as the per-slice kernel below, at a third of its time
        upitch = (rows * cols + 15) & ~15
        stage = torch.empty((ns, upitch), dtype=torch.uint8, device=vol.device)
        rc = lib.msl_stage_slices(_ptr(vol), nvol, X, Y, Z, L.PLANO_ID[plano], _ptr(vs), _ptr(ix), ns, _ptr(stage), upitch, _stream())
        if rc == L.ERR_UNSUPPORTED:
            rc = lib.msl_enhance_slices(
                _ptr(vol), _dtype_id(vol, "vol"), nvol, X, Y, Z, L.MEJORA_NONE, L.PLANO_ID[plano],
                _ptr(vs), _ptr(ix), ns, _ptr(stage), upitch, _LAYOUT_ID["P"], _ptr(tabs), _stream())
        L.check(rc)
        ws = torch.empty(int(lib.msl_enhance_stack_workspace_bytes(rows, cols)) + 16, dtype=torch.uint8, device=vol.device)
        dst = [_ptr(out) if mejora == m else None for m in ("HE", "CLAHE", "GC", "LT")]
        rc = lib.msl_enhance_stack(_ptr(stage), upitch, ns, rows, cols, *dst, _ptr(tabs), _ptr(ws), ws.numel(), _stream())
        if rc == 0:
            return out
        if rc != L.ERR_UNSUPPORTED:
            L.check(rc)
    pitch = rows * cols * (4 if layout == "PNG_RGBA" else 1)
    L.check(lib.msl_enhance_slices(
        _ptr(vol), _dtype_id(vol, "vol"), nvol, X, Y, Z, L.MEJORA_ID[mejora], L.PLANO_ID[plano],
        _ptr(vs), _ptr(ix), ns, _ptr(out), pitch, _LAYOUT_ID[layout], _ptr(tabs), _stream()))
    return out


@_nvtx("enhance_images")
def enhance_images(imgs: torch.Tensor, mejora: Optional[str], layout: str = "G", lut_out: str = "gray") -> torch.Tensor:
    """Batch of C-contiguous 2-D images [n, rows, cols] (float32 or uint8) -> enhanced gray images."""
    _need_cuda(imgs, "imgs")
    if imgs.dim() != 3:
        raise ValueError("imgs must be [n, rows, cols]")
    if mejora not in L.MEJORA_ID:
        raise ValueError(f"Mejora no reconocida: {mejora}.")
    n, rows, cols = imgs.shape
    out = torch.empty(_out_shape(n, rows, cols, layout), dtype=torch.uint8, device=imgs.device)
    pitch = rows * cols * (4 if layout == "PNG_RGBA" else 1)
    L.check(L.load().msl_enhance_images(
        _ptr(imgs), _dtype_id(imgs, "imgs"), n, rows, cols, rows * cols, L.MEJORA_ID[mejora],
        _ptr(out), pitch, _LAYOUT_ID[layout], _ptr(device_tables(imgs.device, lut_out)), _stream()))
    return out


def enhance_volumes_workspace_bytes(nvol: int, X: int, Y: int, Z: int) -> int:
    return int(L.load().msl_workspace_bytes(L.WS_ENHANCE_VOLUMES, nvol, X, Y, Z))


@_nvtx("enhance_volumes")
def enhance_volumes(vol: torch.Tensor, mejoras: Iterable[str] = MEJORAS, planos: Iterable[str] = PLANOS,
                    outs: Optional[Dict[Tuple[str, str], torch.Tensor]] = None,
                    workspace: Optional[torch.Tensor] = None,
                    tables: Optional[torch.Tensor] = None) -> Dict[Tuple[str, str], torch.Tensor]:
    """All slices of the requested planes and enhancements from ONE resident float32 copy.
    Returns {(mejora, plano): uint8 [nvol, n_plane, cols, rows]} in PNG orientation; the slice-oriented
    view is `P.flip(-2).transpose(-1, -2)`.  `tables`: a device copy of a custom MSL_TABLES_BYTES block
    (default: the reference's constants, `device_tables`)."""
    _need_cuda(vol, "vol")
    if vol.dim() != 4 or vol.dtype != torch.float32:
        raise ValueError("vol must be float32 [nvol, Z, Y, X]")
    nvol, Z, Y, X = vol.shape
    mejoras, planos = tuple(mejoras), tuple(planos)
    result: Dict[Tuple[str, str], torch.Tensor] = {}
    ptrs = (C.c_void_p * 12)()
    for m in mejoras:
        if m not in MEJORAS:
            raise ValueError(f"Mejora no reconocida: {m}.")
        for pl in planos:
            n_p, rows, cols = plane_dims(pl, X, Y, Z)
            shape = (nvol, n_p, cols, rows)
            t = None if outs is None else outs.get((m, pl))
            if t is None:
                t = torch.empty(shape, dtype=torch.uint8, device=vol.device)
            else:
                _need_cuda(t, f"outs[{m},{pl}]")
                if tuple(t.shape) != shape or t.dtype != torch.uint8:
                    raise ValueError(f"outs[{m},{pl}] must be uint8 {shape}")
            result[(m, pl)] = t
            ptrs[(L.MEJORA_ID[m] - 1) * 3 + L.PLANO_ID[pl]] = t.data_ptr()
    need = enhance_volumes_workspace_bytes(nvol, X, Y, Z)
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=vol.device)
    _need_cuda(workspace, "workspace")
    if tables is None:
        tables = device_tables(vol.device)
    _need_cuda(tables, "tables")
    if tables.dtype != torch.uint8 or tables.numel() != T.TABLES_BYTES:
        raise ValueError(f"tables must be uint8 [{T.TABLES_BYTES}]")
    L.check(L.load().msl_enhance_volumes(_ptr(vol), nvol, X, Y, Z, ptrs, _ptr(tables),
                                         _ptr(workspace), workspace.numel() * workspace.element_size(), _stream()))
    return result


# ------------------------------------------------------------------------------------ R0
@_nvtx("combine_predictions")
def combine_predictions(masks: Optional[torch.Tensor], inst_offset, rows: int, cols: int, layout: str = "G",
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """YOLO instance masks -> one predicted mask per slice (reference scripts/generar_predicciones.py:123-140).
    masks: float32 [n_inst, mh, mw] (`masks.data` of all slices, concatenated; may be empty), inst_offset: [n + 1]
    prefix offsets.  layout "G": uint8 [n, rows, cols] {0, 255} (normalizar_prediccion, the input of `recon`);
    layout "P": uint8 [n, cols, rows] {0, 1} (combinar_predicciones on the PNG-oriented image)."""
    if layout not in ("G", "P"):
        raise ValueError("layout must be 'G' or 'P'")
    if masks is None or masks.numel() == 0:
        dev = out.device if out is not None else torch.device("cuda", torch.cuda.current_device())
        masks_ptr, mh, mw = None, 1, 1
    else:
        _need_cuda(masks, "masks")
        if masks.dtype != torch.float32 or masks.dim() != 3:
            raise ValueError("masks must be float32 [n_inst, mh, mw]")
        dev, masks_ptr, mh, mw = masks.device, _ptr(masks), int(masks.shape[1]), int(masks.shape[2])
    n_inst = 0 if masks is None else int(masks.shape[0])
    if not isinstance(inst_offset, torch.Tensor):
        # a host table is validated on the host (no synchronisation); a table that already lives on the device is
        # trusted - checking it would cost a blocking device-to-host copy on every call
        host_off = [int(v) for v in np.asarray(inst_offset).reshape(-1)]
        if not host_off:
            raise ValueError("inst_offset needs n + 1 entries")
        if host_off[0] != 0 or host_off[-1] != n_inst or any(b < a for a, b in zip(host_off, host_off[1:])):
            raise ValueError("inst_offset must be a non-decreasing prefix table from 0 to the number of instance masks")
    off = _index_tensor(inst_offset, dev)
    n = int(off.numel()) - 1
    if n < 0:
        raise ValueError("inst_offset needs n + 1 entries")
    shape = (n, rows, cols) if layout == "G" else (n, cols, rows)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=dev)
    _need_cuda(out, "out")
    if tuple(out.shape) != shape or out.dtype != torch.uint8:
        raise ValueError(f"out must be uint8 {shape}")
    L.check(L.load().msl_combine_predictions(masks_ptr, _ptr(off), n, mh, mw, int(rows), int(cols),
                                             L.OUT_G if layout == "G" else L.OUT_P, _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------ R1-R2
@_nvtx("recon")
def recon(slices: torch.Tensor, vol_of_slice, idx_of_slice, plano: str, nvol: int, shape_xyz: Sequence[int],
          dtype: torch.dtype = torch.uint8, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Predicted masks [n, rows, cols] uint8 (pixel > 0 = lesion) -> volumes [nvol, Z, Y, X] of 0/1."""
    X, Y, Z = (int(d) for d in shape_xyz)
    n_p, rows, cols = plane_dims(plano, X, Y, Z)
    _need_cuda(slices, "slices")
    if slices.dtype != torch.uint8 or slices.dim() != 3 or tuple(slices.shape[1:]) != (rows, cols):
        raise ValueError(
            f"Dimensiones {tuple(slices.shape[1:])} incorrectas para plano {plano}. Se esperaba {(rows, cols)}.")
    dev = slices.device
    vs, ix = _index_tensor(vol_of_slice, dev), _index_tensor(idx_of_slice, dev)
    ns = int(slices.shape[0])
    if vs.numel() != ns or ix.numel() != ns:
        raise ValueError("one (volume, index) pair per slice is required")
    if dtype not in (torch.uint8, torch.float32):
        raise TypeError("dtype must be torch.uint8 or torch.float32")
    if out is None:
        out = torch.empty((nvol, Z, Y, X), dtype=dtype, device=dev)
    else:
        _need_cuda(out, "out")
        if tuple(out.shape) != (nvol, Z, Y, X) or out.dtype not in (torch.uint8, torch.float32) or out.device != dev:
            raise ValueError(f"out must be a contiguous uint8 / float32 CUDA tensor {(nvol, Z, Y, X)} on {dev}")
    ws_bytes = int(L.load().msl_workspace_bytes(L.WS_RECON, nvol, X, Y, Z))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    u8 = out if out.dtype == torch.uint8 else None
    f32 = out if out.dtype == torch.float32 else None
    L.check(L.load().msl_recon(_ptr(slices), rows * cols, _ptr(vs), _ptr(ix), ns, L.PLANO_ID[plano], nvol, X, Y, Z,
                               _ptr(u8), _ptr(f32), _ptr(ws), ws_bytes, _stream()))
    return out


# ------------------------------------------------------------------------------------ R3-R4
@_nvtx("consensus_eval")
def consensus_eval(ax: torch.Tensor, co: torch.Tensor, sa: torch.Tensor, gt: Optional[torch.Tensor] = None,
                   umbral: int = 2, want_consenso: bool = True):
    """(ax + co + sa >= umbral) and, when gt is given, int64 counts [nvol, 4 planes, (tp, fp, fn, tn)]
    for axial, coronal, sagital and consenso in the same pass.  Returns (consenso | None, counts | None)."""
    for name, t in (("ax", ax), ("co", co), ("sa", sa)):
        _need_cuda(t, name)
        if t.dtype != torch.uint8 or t.shape != ax.shape:
            raise ValueError("ax / co / sa must be uint8 tensors of one shape")
    nvol = ax.shape[0] if ax.dim() > 1 else 1
    nvox = ax.numel() // nvol
    counts = None
    if gt is not None:
        _need_cuda(gt, "gt")
        if gt.dtype != torch.uint8 or gt.shape != ax.shape:
            raise ValueError("gt must be a uint8 tensor shaped like the plane volumes")
        counts = torch.empty((nvol, 4, 4), dtype=torch.int64, device=ax.device)
    cons = torch.empty_like(ax) if want_consenso else None
    L.check(L.load().msl_consensus_eval(_ptr(ax), _ptr(co), _ptr(sa), _ptr(gt), nvol, nvox, int(umbral),
                                        _ptr(cons), _ptr(counts), _stream()))
    return cons, counts


@_nvtx("confusion_counts")
def confusion_counts(gt: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """int64 [nvol, (tp, fp, fn, tn)] with the reference's exact ==1 / ==0 predicates."""
    _need_cuda(gt, "gt")
    _need_cuda(pred, "pred")
    if gt.dtype != torch.uint8 or pred.dtype != torch.uint8 or gt.shape != pred.shape:
        raise ValueError("gt and pred must be uint8 tensors of one shape")
    nvol = gt.shape[0] if gt.dim() > 1 else 1
    nvox = gt.numel() // nvol
    counts = torch.empty((nvol, 4), dtype=torch.int64, device=gt.device)
    L.check(L.load().msl_confusion_counts(_ptr(gt), _ptr(pred), nvol, nvox, _ptr(counts), _stream()))
    return counts


@_nvtx("slice_counts")
def slice_counts(gt: torch.Tensor, pred: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Per-slice (tp, fp, fn, tn) of every slice of the three planes in one pass (the counts behind
    extras/visualizar_prediccion_corte.py seleccionar_mejor_corte).  gt, pred: uint8 [nvol, Z, Y, X].
    Returns {plano: int64 [nvol, n_plane, 4]}."""
    _need_cuda(gt, "gt")
    _need_cuda(pred, "pred")
    if gt.dtype != torch.uint8 or pred.dtype != torch.uint8 or gt.shape != pred.shape or gt.dim() != 4:
        raise ValueError("gt and pred must be uint8 tensors [nvol, Z, Y, X] of one shape")
    nvol, Z, Y, X = (int(d) for d in gt.shape)
    counts = torch.empty((nvol, Z + Y + X, 4), dtype=torch.int64, device=gt.device)
    L.check(L.load().msl_slice_counts(_ptr(gt), _ptr(pred), nvol, X, Y, Z, _ptr(counts), _stream()))
    return {"axial": counts[:, :Z], "coronal": counts[:, Z:Z + Y], "sagital": counts[:, Z + Y:]}


def bgr_to_gray(bgr: torch.Tensor) -> torch.Tensor:
    """cv2.cvtColor(BGR2GRAY) of uint8 [..., 3] images (reference utils/utils.py:421-427 verificar_grises)."""
    _need_cuda(bgr, "bgr")
    if bgr.dtype != torch.uint8 or bgr.dim() < 1 or bgr.shape[-1] != 3:
        raise ValueError("bgr must be uint8 [..., 3]")
    bgr = bgr.contiguous()
    gray = torch.empty(bgr.shape[:-1], dtype=torch.uint8, device=bgr.device)
    L.check(L.load().msl_bgr_to_gray(_ptr(bgr), gray.numel(), _ptr(gray), _stream()))
    return gray


def png_pack(pixels: torch.Tensor):
    """Complete PNG files for a batch of images: uint8 [n, H, W, 4] (RGBA, what plt.imsave writes; reference
    scripts/extraer_dataset.py:192,197) or [n, H, W] (gray).  Returns (files, size): uint8 [n, pitch] on the device and
    the length of each file; `files[i, :size].cpu().numpy().tobytes()` is file i (stored deflate blocks)."""
    _need_cuda(pixels, "pixels")
    if pixels.dtype != torch.uint8 or pixels.dim() not in (3, 4) or (pixels.dim() == 4 and pixels.shape[-1] != 4):
        raise ValueError("pixels must be uint8 [n, H, W, 4] or [n, H, W]")
    pixels = pixels.contiguous()
    n, H, W = (int(d) for d in pixels.shape[:3])
    ch = 4 if pixels.dim() == 4 else 1
    size = int(L.load().msl_png_bytes(H, W, ch))
    pitch = (size + 15) & ~15
    out = torch.empty((n, pitch), dtype=torch.uint8, device=pixels.device)
    L.check(L.load().msl_png_pack(_ptr(pixels), n, H, W, ch, _ptr(out), pitch, _stream()))
    return out, size


# ------------------------------------------------------------------------------------ byte-stream codec
_CONTAINER_ID = {"raw": L.Z_RAW, "zlib": L.Z_ZLIB, "gzip": L.Z_GZIP, "png": L.Z_PNG}


class PackedStreams:
    """n variable-length byte streams packed back to back on the device: stream i = data[off[i]:off[i+1]] (off: uint64
    [n + 1], device).  `to_host()` brings offsets and bytes to the host (two copies: the offsets tell how many bytes)."""

    def __init__(self, data: torch.Tensor, off: torch.Tensor, meta: Optional[torch.Tensor] = None):
        self.data, self.off, self.meta = data, off, meta

    def to_host(self, pinned: Optional[torch.Tensor] = None):
        """(bytes ndarray of the packed streams, offsets ndarray int64 [n + 1])."""
        off = self.off.cpu().numpy().astype(np.int64)
        total = int(off[-1])
        if total > self.data.numel():
            raise RuntimeError(f"packed streams need {total} bytes, capacity was {self.data.numel()}")
        if pinned is not None:
            pinned[:total].copy_(self.data[:total], non_blocking=False)
            return pinned[:total].numpy(), off
        return self.data[:total].cpu().numpy(), off

    def files(self):
        data, off = self.to_host()
        return [data[off[i]:off[i + 1]].tobytes() for i in range(len(off) - 1)]


@_nvtx("deflate_chunks")
def deflate_chunks(src: torch.Tensor, chunk_len: int = 65536, container: str = "gzip", dist2: int = 0,
                   out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> PackedStreams:
    """Cuts the bytes of `src` (any contiguous CUDA tensor) into chunks of chunk_len bytes and deflates every chunk into
    its own stream (container "raw", "zlib" or "gzip"; dist2 = the match distance 1..4, 0 = 1: the element size in bytes).  The concatenation of the gzip members is a valid .gz file of the
    whole buffer (reference utils/utils.py:176-177 nib.save -> gzip)."""
    _need_cuda(src, "src")
    lib = L.load()
    cid = _CONTAINER_ID[container]
    total = src.numel() * src.element_size()
    n = max(1, -(-total // chunk_len))
    cap = int(lib.msl_deflate_bound(n, cid, chunk_len))
    if out is None:
        out = torch.empty(cap, dtype=torch.uint8, device=src.device)
    need = int(lib.msl_deflate_workspace_bytes(n, cid, chunk_len))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=src.device)
    off = torch.empty(n + 1, dtype=torch.int64, device=src.device)
    meta = torch.empty((n, 4), dtype=torch.int32, device=src.device)
    L.check(lib.msl_deflate_chunks(_ptr(src), total, chunk_len, cid, dist2, _ptr(out), out.numel(), _ptr(off), _ptr(meta),
                                   _ptr(workspace), workspace.numel(), _stream()))
    return PackedStreams(out, off, meta)


@_nvtx("deflate_files")
def deflate_files(bodies: torch.Tensor, prefix: Optional[torch.Tensor] = None, expand_u8_to_f32: bool = False, chunk_len: int = 65536,
                  container: str = "gzip", dist2: int = 0, out: Optional[torch.Tensor] = None,
                  workspace: Optional[torch.Tensor] = None) -> PackedStreams:
    """nfiles = bodies.shape[0] files in one launch: file f = prefix (one for all, or [nfiles, prefix_len]) + bodies[f]
    (any dtype, contiguous), cut into chunks of chunk_len bytes, every chunk its own deflate stream; streams are numbered
    file-major, `streams_per_file` of them per file.  expand_u8_to_f32: bodies are uint8 masks stored as float32 0 / 1."""
    _need_cuda(bodies, "bodies")
    lib = L.load()
    cid = _CONTAINER_ID[container]
    nfiles = int(bodies.shape[0])
    body_len = (bodies.numel() // max(nfiles, 1)) * bodies.element_size()
    plen, ppitch = 0, 0
    if prefix is not None:
        _need_cuda(prefix, "prefix")
        if prefix.dtype != torch.uint8 or prefix.dim() not in (1, 2) or (prefix.dim() == 2 and prefix.shape[0] != nfiles):
            raise ValueError("prefix must be uint8 [prefix_len] or [nfiles, prefix_len]")
        plen = int(prefix.shape[-1])
        ppitch = plen if prefix.dim() == 2 else 0
    if expand_u8_to_f32 and bodies.dtype != torch.uint8:
        raise ValueError("expand_u8_to_f32 needs uint8 bodies")
    total = plen + body_len * (4 if expand_u8_to_f32 else 1)
    spv = max(1, -(-total // chunk_len))
    n = spv * nfiles
    cap = int(lib.msl_deflate_bound(n, cid, chunk_len))
    if out is None:
        out = torch.empty(cap, dtype=torch.uint8, device=bodies.device)
    need = int(lib.msl_deflate_workspace_bytes(n, cid, chunk_len))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=bodies.device)
    off = torch.empty(n + 1, dtype=torch.int64, device=bodies.device)
    meta = torch.empty((n, 4), dtype=torch.int32, device=bodies.device)
    L.check(lib.msl_deflate_files(_ptr(bodies), nfiles, body_len, body_len, _ptr(prefix), ppitch, plen, 1 if expand_u8_to_f32 else 0,
                                  chunk_len, cid, dist2, _ptr(out), out.numel(), _ptr(off), _ptr(meta), _ptr(workspace),
                                  workspace.numel(), _stream()))
    ps = PackedStreams(out, off, meta)
    ps.streams_per_file = spv
    return ps


@_nvtx("png_encode")
def png_encode(pixels: torch.Tensor, out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> PackedStreams:
    """Compressed PNG files for a batch of images: uint8 [n, H, W, C] (C = 4 RGBA - what plt.imsave writes, reference
    scripts/extraer_dataset.py:192,197 - or 2 / 3) or [n, H, W] (gray, what cv2.imwrite writes for masks).  The files are
    packed back to back on the device (PackedStreams)."""
    _need_cuda(pixels, "pixels")
    if pixels.dtype != torch.uint8 or pixels.dim() not in (3, 4) or (pixels.dim() == 4 and not 1 <= pixels.shape[-1] <= 4):
        raise ValueError("pixels must be uint8 [n, H, W, C] (C <= 4) or [n, H, W]")
    pixels = pixels.contiguous()
    n, H, W = (int(d) for d in pixels.shape[:3])
    ch = int(pixels.shape[3]) if pixels.dim() == 4 else 1
    lib = L.load()
    raw = H * (W * ch + 1)
    cap = int(lib.msl_deflate_bound(max(n, 1), L.Z_PNG, raw))
    if out is None:
        out = torch.empty(cap, dtype=torch.uint8, device=pixels.device)
    need = int(lib.msl_deflate_workspace_bytes(max(n, 1), L.Z_PNG, raw))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=pixels.device)
    off = torch.zeros(n + 1, dtype=torch.int64, device=pixels.device)
    L.check(lib.msl_png_encode(_ptr(pixels), n, H, W, ch, _ptr(out), out.numel(), _ptr(off), _ptr(workspace), workspace.numel(), _stream()))
    return PackedStreams(out, off)


# ------------------------------------------------------------------------------------ label polygons
@_nvtx("mask_contours")
def mask_contours(masks: torch.Tensor, value: int = 0, max_contours: int = 256, max_points: int = 8192):
    """External contours of binary masks uint8 [n, H, W], as cv2.findContours(mask == value (or mask != 0 when value is
    0), RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) returns them (reference scripts/extraer_dataset.py:215-227 via ultralytics).
    Returns a list (one entry per mask) of lists of int32 arrays [npoints, 2] (x, y) in OpenCV's order.  Capacities are
    grown and the call repeated when a mask holds more contours / points than asked for."""
    _need_cuda(masks, "masks")
    if masks.dtype != torch.uint8 or masks.dim() != 3:
        raise ValueError("masks must be uint8 [n, H, W]")
    masks = masks.contiguous()
    n, H, W = (int(d) for d in masks.shape)
    if n == 0:
        return []
    while True:
        counts = torch.empty((n, 4), dtype=torch.int32, device=masks.device)
        clen = torch.empty((n, max_contours), dtype=torch.int32, device=masks.device)
        pts = torch.empty((n, max_points, 2), dtype=torch.int16, device=masks.device)
        L.check(L.load().msl_mask_contours(_ptr(masks), n, H, W, int(value), max_contours, max_points, _ptr(counts), _ptr(clen),
                                           _ptr(pts), _stream()))
        c = counts.cpu().numpy()
        if (c[:, 2] & 1).any():
            max_contours = int(c[:, 3].max()) + 1
            continue
        if (c[:, 2] & 2).any():
            max_points *= 4
            continue
        break
    ncmax, npmax = int(c[:, 0].max()), int(c[:, 1].max())
    lens = clen[:, :max(ncmax, 1)].cpu().numpy()
    pp = pts[:, :max(npmax, 1)].cpu().numpy().astype(np.int32)
    out = []
    for i in range(n):
        o, cs = 0, []
        for k in range(int(c[i, 0])):
            cs.append(pp[i, o:o + int(lens[i, k])])
            o += int(lens[i, k])
        out.append(cs)
    return out


# ------------------------------------------------------------------------------------ host hand-off
@_nvtx("nonzero_flags")
def nonzero_flags(stack: torch.Tensor, out=None):
    """Which slices and rows of a uint8 stack [nvol, A, B, C] hold a non-zero byte: (any_a [nvol, A], any_b [nvol, B]).
    out: optional pair of contiguous uint8 device tensors of those shapes (e.g. views of one buffer that goes to the host
    in a single copy)."""
    _need_cuda(stack, "stack")
    if stack.dtype != torch.uint8 or stack.dim() != 4 or not stack.is_contiguous():
        raise ValueError("stack must be a contiguous uint8 tensor [nvol, A, B, C]")
    nvol, A, B, Cc = (int(d) for d in stack.shape)
    if out is None:
        any_a = torch.empty((nvol, A), dtype=torch.uint8, device=stack.device)
        any_b = torch.empty((nvol, B), dtype=torch.uint8, device=stack.device)
    else:
        any_a, any_b = out
        _need_cuda(any_a, "out[0]")
        _need_cuda(any_b, "out[1]")
        if tuple(any_a.shape) != (nvol, A) or tuple(any_b.shape) != (nvol, B) or any_a.dtype != torch.uint8 or any_b.dtype != torch.uint8:
            raise ValueError("out must be uint8 tensors [nvol, A] and [nvol, B]")
    L.check(L.load().msl_nonzero_flags(_ptr(stack), nvol, A, B, Cc, _ptr(any_a), _ptr(any_b), _stream()))
    return any_a, any_b


def box_from_flags(any_a: np.ndarray, any_b: np.ndarray):
    """(a0, a1, b0, b1) of one volume from its host-side flags; an empty box is (0, 0, 0, 0)."""
    ia, ib = np.flatnonzero(any_a), np.flatnonzero(any_b)
    if ia.size == 0 or ib.size == 0:
        return (0, 0, 0, 0)
    return (int(ia[0]), int(ia[-1]) + 1, int(ib[0]), int(ib[-1]) + 1)


def copy_box_to_host(dev: torch.Tensor, host: torch.Tensor, box) -> int:
    """Asynchronous device-to-host copy (current stream) of the box (a0, a1, b0, b1) - full rows - of ONE uint8 array
    [A, B, C] into a (pinned) host array of the same shape; the rest of `host` is left alone (the caller keeps it zero).
    Returns the number of bytes copied."""
    _need_cuda(dev, "dev")
    if dev.dtype != torch.uint8 or dev.dim() != 3 or not dev.is_contiguous():
        raise ValueError("dev must be a contiguous uint8 tensor [A, B, C]")
    if host.device.type != "cpu" or host.dtype != torch.uint8 or tuple(host.shape) != tuple(dev.shape) or not host.is_contiguous():
        raise ValueError("host must be a contiguous uint8 CPU tensor of the same shape")
    A, B, Cc = (int(d) for d in dev.shape)
    a0, a1, b0, b1 = (int(v) for v in box)
    L.check(L.load().msl_copy_box_d2h(host.data_ptr(), _ptr(dev), A, B, Cc, a0, a1, b0, b1, _stream()))
    return (a1 - a0) * (b1 - b0) * Cc


class HostResult:
    """A pinned host array [nvol, A, B, C] that mirrors device results while receiving only their non-zero boxes.
    Successive updates of the same volumes may be in flight together; only when a box shrinks (stale data must be zeroed
    on the host) does `update` wait for the earlier copy of that volume."""

    def __init__(self, shape):
        self.host = torch.zeros(tuple(int(d) for d in shape), dtype=torch.uint8).pin_memory()
        self.boxes = [(0, 0, 0, 0)] * int(shape[0])
        self.events = [None] * int(shape[0])

    def update(self, v0: int, dev_stack: torch.Tensor, any_a: np.ndarray, any_b: np.ndarray) -> int:
        """Volumes v0 .. v0+n of the host array := dev_stack [n, A, B, C], given its host-side non-zero flags.
        Copies are enqueued on the current stream; returns the bytes they move."""
        _need_cuda(dev_stack, "dev_stack")
        n = int(dev_stack.shape[0])
        if dev_stack.dtype != torch.uint8 or tuple(dev_stack.shape[1:]) != tuple(self.host.shape[1:]) or v0 < 0 or v0 + n > self.host.shape[0]:
            raise ValueError("dev_stack must be uint8 [n, A, B, C] and fit the host array")
        A, B, Cc = (int(d) for d in self.host.shape[1:])
        boxes = np.zeros((n, 4), dtype=np.int32)
        moved = 0
        for i in range(n):
            box = box_from_flags(any_a[i], any_b[i])
            old = self.boxes[v0 + i]
            if old[1] > old[0] and not (box[0] <= old[0] and old[1] <= box[1] and box[2] <= old[2] and old[3] <= box[3]):
                if self.events[v0 + i] is not None:
                    self.events[v0 + i].synchronize()                        # the earlier copy into this volume has landed
                self.host[v0 + i, old[0]:old[1], old[2]:old[3]] = 0          # stale non-zero data outside the new box
            self.boxes[v0 + i] = box
            boxes[i] = box
            moved += (box[1] - box[0]) * (box[3] - box[2]) * Cc
        L.check(L.load().msl_copy_boxes_d2h(self.host[v0].data_ptr(), _ptr(dev_stack), n, A, B, Cc,
                                            boxes.ctypes.data_as(C.c_void_p), _stream()))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        for i in range(n):
            self.events[v0 + i] = ev
        return moved
