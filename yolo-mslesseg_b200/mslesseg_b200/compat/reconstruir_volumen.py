"""Mirror of the numeric part of yolo_mslesseg/scripts/reconstruir_volumen.py (:108-213)."""
from __future__ import annotations

import logging
import re
from pathlib import Path

import numpy as np
import torch
from PIL import Image

from .. import codec as _codec
from .. import ops
from . import device
from .utils import cargar_referencia_nifti, guardar_volumen, ruta_existente

logger = logging.getLogger(__name__)


def extraer_indices_png(input_dir):
    input_dir = Path(input_dir)
    if not ruta_existente(input_dir):
        raise FileNotFoundError(f"No se encontró el directorio de máscaras predichas: {input_dir}")
    patron = re.compile(r".*_(\d+)(?:_[^_]*)?\.png$", re.IGNORECASE)
    tuplas = []
    for p in input_dir.glob("*.png"):
        m = patron.match(p.name)
        if m:
            tuplas.append((p.name, int(m.group(1))))
        else:
            logger.warning(f"⚠️ No se pudo extraer el índice de {p.name}")
    if not tuplas:
        raise FileNotFoundError("No hay máscaras predichas que procesar.")
    tuplas.sort(key=lambda t: t[1])
    return tuplas


def cargar_mascara_png(img_path) -> np.ndarray:
    """PIL decode + channel-0 selection of cargar_y_preprocesar_imagen (:136-145); the binarisation (:146-148)
    happens on the GPU (voxel = pixel > 0, identical for {0,255} and {0,1} masks)."""
    if not ruta_existente(img_path):
        raise FileNotFoundError(f"No se encontró la imagen: {img_path}")
    a = np.array(Image.open(img_path))
    if a.ndim > 2:
        a = a[:, :, 0]
    if a.dtype != np.uint8:
        a = (a > 0).astype(np.uint8)
    return a


def validar_corte(indice, img_array, shape_original, plano):
    max_indices = {"axial": shape_original[2], "coronal": shape_original[1], "sagital": shape_original[0]}
    if indice < 0 or indice >= max_indices[plano]:
        raise ValueError(f"Índice {indice} fuera de rango para plano {plano}.")
    expected = {"axial": (shape_original[0], shape_original[1]), "coronal": (shape_original[0], shape_original[2]),
                "sagital": (shape_original[1], shape_original[2])}[plano]
    if img_array.shape != expected:
        raise ValueError(f"Dimensiones {img_array.shape} incorrectas para plano {plano}. Se esperaba {expected}.")


def reconstruir_desde_cortes(cortes, indices, shape_original, plano) -> np.ndarray:
    """In-memory core: list of 2-D masks + indices -> float32 (X, Y, Z) volume (zeros where no slice exists)."""
    for q, i in zip(cortes, indices):
        validar_corte(int(i), np.asarray(q), shape_original, plano)
    X, Y, Z = (int(d) for d in shape_original)
    n_p, rows, cols = ops.plane_dims(plano, X, Y, Z)
    if len(cortes):
        st = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(q, dtype=np.uint8) for q in cortes]))).to(device())
    else:
        st = torch.zeros((0, rows, cols), dtype=torch.uint8, device=device())
    vol = ops.recon(st, [0] * len(cortes), [int(i) for i in indices], plano, 1, (X, Y, Z), dtype=torch.float32)
    return np.asfortranarray(vol[0].cpu().numpy().transpose(2, 1, 0))


def reconstruir_volumen(pred_masks_dir, volumen_referencia, output_path, plano):
    """Same signature and side effects as the reference (:199-213): writes a float32 NIfTI, returns the volume.
    The predicted-mask PNG FILES are uploaded as they are: inflate, scanline unfiltering, channel-0 selection and the
    binarisation run on the GPU, the volume is stacked there and its .nii.gz is deflated there."""
    shape_original, affine = cargar_referencia_nifti(volumen_referencia)
    indices = extraer_indices_png(pred_masks_dir)
    archivos = []
    for archivo, _ in indices:
        ruta = Path(pred_masks_dir) / archivo
        if not ruta_existente(ruta):
            raise FileNotFoundError(f"No se encontró la imagen: {ruta}")
        archivos.append(ruta.read_bytes())
    X, Y, Z = (int(d) for d in shape_original)
    n_p, rows, cols = ops.plane_dims(plano, X, Y, Z)
    geos = [_codec.png_parse(b)[:2] for b in archivos]                       # (width, height) of every file
    for (archivo, i), (w, h) in zip(indices, geos):
        validar_corte(int(i), np.empty((h, w), dtype=np.uint8), shape_original, plano)   # index range + exact slice shape (:153-176)
    st = _codec.png_decode_first_channel(archivos, device())                 # [n, rows, cols] uint8, slice orientation
    vol = ops.recon(st, [0] * len(indices), [int(i) for _, i in indices], plano, 1, (X, Y, Z))        # uint8 {0, 1} on the device
    try:
        blob = _codec.nifti_gz_bytes(_codec.nifti_gz_device(vol[0], affine, como_float32=True))     # the file stores float32 (:202)
        Path(output_path).parent.mkdir(parents=True, exist_ok=True)
        with open(output_path, "wb") as f:
            f.write(blob)
    except Exception as e:
        logger.error(f"❌ Error al guardar el volumen en {output_path}: {e}")
        raise
    return np.asfortranarray(vol[0].cpu().numpy().transpose(2, 1, 0).astype(np.float32))
